set -x
mkdir -p gpurun_out
timeout 300 python tools/tfam_fused_debug.py > gpurun_out/r02_tfam_debug2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tfam_debug2.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_pytest2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench2_default.json 2> gpurun_out/r02_bench2_default.err
tail -n 5 gpurun_out/r02_pytest2.log
tail -n 6 gpurun_out/r02_tfam_debug2.log
