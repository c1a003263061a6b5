set -x
mkdir -p gpurun_out
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bw tools/micro/mufu_bw.cu && /tmp/mufu_bw) > gpurun_out/r02_mufu2.log 2>&1
timeout 600 python tools/tower_margins.py > gpurun_out/r02_tower_margins.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "scale or resize or student_config1 or tfam_fused" > gpurun_out/r02_pytest3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest3.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench3_default.json 2> gpurun_out/r02_bench3_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench3_reference.json 2> gpurun_out/r02_bench3_reference.err
tail -n 4 gpurun_out/r02_pytest3.log; cat gpurun_out/r02_mufu2.log | tail -n 6; cat gpurun_out/r02_tower_margins.log
