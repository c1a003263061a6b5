set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_pytest_final2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_final2.log
tail -n 4 gpurun_out/r02_pytest_final2.log
python bench.py --steps 5 --warmup 3 --no-configs --no-strong --no-ab > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench8.json
