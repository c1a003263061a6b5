set -x
mkdir -p gpurun_out
python tools/ncu_probe_r02.py > gpurun_out/r02_probe_plain2.log 2>&1; echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:gemm2_bf16|attention_vit5|attention_vit7' -s 6 -c 6 -f -o gpurun_out/r02_final_kernels python tools/ncu_probe_r02.py > gpurun_out/r02_probe_ncu3.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_final_kernels.ncu-rep
