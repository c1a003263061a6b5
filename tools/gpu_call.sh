set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "gemm or tower or pipeline or config or tfam or student_config1" > gpurun_out/r02_pytest10.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest10.log
true
python - > gpurun_out/r02_gemm_times2.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from vimoclip_b200 import ops
dev = torch.device("cuda:0"); gen = torch.Generator(device="cuda").manual_seed(0)
F_, L, d = 2048, 197, 768
M = F_ * L
xs = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
parts = ops.gemm_stats_parts(M, d); stats = torch.zeros(parts, M, 2, device=dev)
rnd = lambda n, k: (torch.randn(n, k, device=dev, generator=gen) * k**-0.5).to(torch.bfloat16)
def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
a = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
w_o, b_o = rnd(d, d), torch.randn(d, device=dev)
ops.gemm(a, w_o, bias=b_o, resid=xs, out=xs, emit_stats=(None, stats))
w_q, b_q = rnd(3 * d, d), torch.randn(3 * d, device=dev); cs_q = w_q.float().sum(1)
w_f, b_f = rnd(4 * d, d), torch.randn(4 * d, device=dev); cs_f = w_f.float().sum(1)
w_p, b_p = rnd(d, 4 * d), torch.randn(d, device=dev)
qkv = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16); h = torch.empty(M, 4 * d, device=dev, dtype=torch.bfloat16)
for name, fn, fl in [
    ("out_proj bf16-resid+stats", lambda: ops.gemm(a, w_o, bias=b_o, resid=xs, out=xs, emit_stats=(None, stats)), 2.0 * M * d * d),
    ("qkv LN-fold", lambda: ops.gemm(xs, w_q, bias=b_q, out=qkv, fold=(stats, cs_q, 1e-5)), 2.0 * M * 3 * d * d),
    ("qkv plain", lambda: ops.gemm(xs, w_q, bias=b_q, out=qkv), 2.0 * M * 3 * d * d),
    ("c_fc LN-fold+QuickGELU", lambda: ops.gemm(xs, w_f, bias=b_f, act=ops.ACT_QUICKGELU, out=h, fold=(stats, cs_f, 1e-5)), 2.0 * M * 4 * d * d),
    ("c_fc plain+QuickGELU", lambda: ops.gemm(xs, w_f, bias=b_f, act=ops.ACT_QUICKGELU, out=h), 2.0 * M * 4 * d * d),
    ("c_proj bf16-resid+stats", lambda: ops.gemm(h, w_p, bias=b_p, resid=xs, out=xs, emit_stats=(None, stats)), 2.0 * M * 4 * d * d),
]:
    ms = timeit(fn)
    print(f"{name:28s} M={M}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
PY
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-strong --no-ab > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err
tail -n 3 gpurun_out/r02_pytest10.log; cat gpurun_out/r02_gemm_times2.log
