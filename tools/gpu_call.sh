set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "ddp or dataparallel or sharded or training or train" > gpurun_out/r02_pytest8_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest8_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
export CUDA_VISIBLE_DEVICES=0
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-strong --no-ab"
K='regex:gemm2_bf16|gemm_bf16|attention_|layernorm_kernel|prologue_|tfam_|student_heads|frame_diff|cast_bf16|mean_rows'
$B > gpurun_out/r02_launchlist_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 275 -c 275 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launchlist_ncu.log 2>&1
python tools/ncu_probe_r02.py > gpurun_out/r02_probe_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16" -s 4 -c 4 -o gpurun_out/r02_gemm_kernels python tools/ncu_probe_r02.py > gpurun_out/r02_probe_ncu2.log 2>&1
tail -n 6 gpurun_out/r02_pytest8_2gpu.log; tail -c 600 gpurun_out/r02_bench_n2.err
