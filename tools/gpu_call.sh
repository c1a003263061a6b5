set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "gemm or tower or vit_" > gpurun_out/r02_pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gemm.log
tail -n 3 gpurun_out/r02_pytest_gemm.log
timeout 300 python tools/gemm_timeline.py > gpurun_out/r02_gemm_timeline2.txt 2>&1
grep "==\|per tile" gpurun_out/r02_gemm_timeline2.txt
GS_ONLY=out_proj,c_proj,qkv timeout 300 python tools/gemm_shapes.py > gpurun_out/r02_gemm_shapes2.log 2>&1
cat gpurun_out/r02_gemm_shapes2.log
