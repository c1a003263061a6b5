set -x
mkdir -p gpurun_out
KB_WHICH=attn KB_ATTN_IMPLS=5,9,6,2 python tools/kernel_bench.py > gpurun_out/r02_kb_attn.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_pytest4.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest4.log
timeout 600 python tools/tower_margins.py > gpurun_out/r02_tower_margins2.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-strong --no-ab > gpurun_out/r02_bench4_attn5.json 2> gpurun_out/r02_bench4_attn5.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-strong --no-ab --attn-impl 9 > gpurun_out/r02_bench4_attn9.json 2> gpurun_out/r02_bench4_attn9.err
cat gpurun_out/r02_kb_attn.log; tail -n 5 gpurun_out/r02_pytest4.log; cat gpurun_out/r02_tower_margins2.log
