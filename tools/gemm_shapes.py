"""Sustained (power-capped) throughput of the tower's GEMMs in the epilogue modes the default path uses, ViT-B/16 at
2048 frames in flight (M = 403 456), plus variants that isolate one epilogue ingredient (no activation, no fold).
Each case runs back to back for GS_SECONDS; CUDA events around the whole run; the SM clock is sampled through NVML
while the case runs (the step is power-capped, so TFLOP/s and clock are reported together).  For tuning, not bench.py."""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vimoclip_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M, d = int(os.environ.get("GS_FRAMES", "2048")) * 197, 768
SECONDS = float(os.environ.get("GS_SECONDS", "1.5"))
ONLY = [s for s in os.environ.get("GS_ONLY", "").split(",") if s]
gen = torch.Generator(device="cuda").manual_seed(0)

try:
    import pynvml

    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:  # noqa: BLE001
    _h = None


class Clocks:
    def __enter__(self):
        self.v, self.p, self.stop = [], [], False

        def run():
            while not self.stop and _h is not None:
                self.v.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM))
                self.p.append(pynvml.nvmlDeviceGetPowerUsage(_h) / 1000.0)
                time.sleep(0.05)

        self.t = threading.Thread(target=run)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join()

    def text(self):
        if not self.v:
            return "clock n/a"
        v = sorted(self.v[len(self.v) // 3:])  # skip the ramp
        p = sorted(self.p[len(self.p) // 3:])
        return f"{v[len(v) // 2]} MHz {p[len(p) // 2]:.0f} W"


def sustained(fn, seconds=SECONDS):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    with Clocks() as c:
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(10):
                fn()
            n += 10
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, c.text()


x = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)
h = torch.randn(M, 4 * d, device=dev, generator=gen).to(torch.bfloat16)
xs = torch.randn(M, d, device=dev, generator=gen).to(torch.bfloat16)  # bf16 residual stream
parts = ops.gemm_stats_parts(M, d)
stats = torch.zeros(parts, M, 2, device=dev)
# plausible row statistics for the fold consumers (sum, sum of squares of x's rows)
xf = x.float()
stats_in = torch.zeros(1, M, 2, device=dev)
stats_in[0, :, 0] = xf.sum(1)
stats_in[0, :, 1] = (xf * xf).sum(1)
del xf


IMPLS = [int(v) for v in os.environ.get("GS_IMPLS", "0").split(",")]  # VMC_OPT_GEMM_IMPL values to time side by side (3 = LDS + STG epilogue)
import vimoclip_b200 as vmc  # noqa: E402


def case(name, a, N, K, **kw):
    if ONLY and not any(s in name for s in ONLY):
        return
    if len(IMPLS) > 1 or IMPLS[0] != 0:
        for impl in IMPLS:
            ops.set_option(vmc._lib.OPT_GEMM_IMPL, impl)
            _case(f"{name} [impl {impl}]", a, N, K, **dict(kw))
        ops.set_option(vmc._lib.OPT_GEMM_IMPL, 0)
        return
    _case(name, a, N, K, **kw)


def _case(name, a, N, K, **kw):
    w = (torch.randn(N, K, device=dev, generator=gen) * K**-0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev, generator=gen)
    if kw.pop("fold", False):
        kw["fold"] = (stats_in, torch.randn(N, device=dev, generator=gen), 1e-5)
    out = kw.pop("out", None)
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ms, clk = sustained(lambda: ops.gemm(a, w, bias=b, out=out, **kw))
    print(f"{name:38s} M={M} N={N:4d} K={K:4d}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:5.0f} TFLOP/s  [{clk}]", flush=True)


case("qkv LN-fold", x, 3 * d, d, fold=True)
case("qkv plain", x, 3 * d, d)
case("c_fc LN-fold+QuickGELU", x, 4 * d, d, fold=True, act=ops.ACT_QUICKGELU)
case("c_fc LN-fold (no act)", x, 4 * d, d, fold=True)
case("c_fc plain+QuickGELU", x, 4 * d, d, act=ops.ACT_QUICKGELU)
case("c_fc plain (no act)", x, 4 * d, d)
case("out_proj shape, plain bias", x, d, d)
case("out_proj bf16-resid+stats", x, d, d, resid=xs, out=xs, emit_stats=(None, stats))
case("c_proj bf16-resid+stats", h, d, 4 * d, resid=xs, out=xs, emit_stats=(None, stats))
if os.environ.get("GS_CUBLAS", "0") == "1":
    for name, a, N, K in [("cuBLAS qkv", x, 3 * d, d), ("cuBLAS c_fc", x, 4 * d, d), ("cuBLAS out_proj", x, d, d), ("cuBLAS c_proj", h, d, 4 * d)]:
        w = (torch.randn(N, K, device=dev, generator=gen) * K**-0.5).to(torch.bfloat16)
        o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms, clk = sustained(lambda: torch.matmul(a, w.t(), out=o2))
        print(f"{name:38s} M={M} N={N:4d} K={K:4d}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:5.0f} TFLOP/s  [{clk}]", flush=True)
