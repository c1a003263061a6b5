set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "attention_vit or tower or pipeline or config" > gpurun_out/r02_pytest7.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest7.log
KB_WHICH=attn KB_ATTN_IMPLS=5,6,7 timeout 300 python tools/kernel_bench.py > gpurun_out/r02_kb_attn3.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-strong --no-ab > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
tail -n 4 gpurun_out/r02_pytest7.log; cat gpurun_out/r02_kb_attn3.log
