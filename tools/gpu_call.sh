set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-configs --no-ab --strong-clips 32 > gpurun_out/r02_bench_strong32.json 2> gpurun_out/r02_bench_strong32.err; echo "rc=$?"
python bench.py --steps 10 --warmup 3 --no-configs --no-ab --no-strong --clips 32 > gpurun_out/r02_bench_clips32.json 2> gpurun_out/r02_bench_clips32.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_strong32.json').read().strip().split('\n')[-1])
print('strong32', d.get('strong'))
d=json.loads(open('gpurun_out/r02_bench_clips32.json').read().strip().split('\n')[-1])
print('clips32', d['ms_per_step'], d['value'], d['gpu_launches'], d['clocks'], d['e2e'])
print({k:(round(v['ms'],2), v['launches']) for k,v in d['kernel_classes'].items()})
PY
