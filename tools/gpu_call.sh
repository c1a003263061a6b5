set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02_pytest_final4.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_final4.log
tail -n 4 gpurun_out/r02_pytest_final4.log
python bench.py --steps 5 --warmup 3 --no-configs --no-strong > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench10.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d.get('ab_same_process'))
print({k:round(v['ms'],2) for k,v in d['kernel_classes'].items()})
PY
