set -x
mkdir -p gpurun_out
touch vimo-clip_b200/csrc/attention.cu
make -C vimo-clip_b200/csrc -j8 > gpurun_out/r02_build.log 2>&1   # rebuild without WHATIF in case the tree travelled with it
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_pytest_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_final.log
tail -n 5 gpurun_out/r02_pytest_final.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -n 2 gpurun_out/r02_smoke.log
tail -c 600 gpurun_out/r02_bench_final_n1.json; tail -c 400 gpurun_out/r02_bench_reference_arm.json
