set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-strong --no-ab"
$B > gpurun_out/r02_launchlist_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 275 -c 275 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launchlist_ncu.log 2>&1
python tools/ncu_probe_r02.py > gpurun_out/r02_probe_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16|attention_vit|tfam_fused" -s 10 -c 10 -o gpurun_out/r02_hot_kernels python tools/ncu_probe_r02.py > gpurun_out/r02_probe_ncu.log 2>&1
ls -la gpurun_out/ | tail -5
