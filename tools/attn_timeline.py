"""Phase timeline (clock64) of CTA 0 of attention_vit5_kernel (impl 55: stamped instantiation; 54: stamped, no softmax)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc
from vimoclip_b200 import ops
dev = torch.device("cuda:0")
F_, L, heads = 1024, 197, 12
impl = int(os.environ.get("TL_IMPL", "55"))
import vimoclip_b200 as vmc  # noqa: E402
ops.set_option(vmc._lib.OPT_ATTN_PREFETCH, int(os.environ.get("KB_ATTN_PREFETCH", "0")))
qkv = torch.randn(F_ * L, 3 * heads * 64, device=dev).to(torch.bfloat16)
dbg = torch.zeros(16 * 32, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.attention_vit(qkv, F_, L, heads, impl=impl)
ops.set_option(7, dbg.data_ptr())
ops.attention_vit(qkv, F_, L, heads, impl=impl)
torch.cuda.synchronize()
ops.set_option(7, 0)
d = dbg.cpu().view(16, 32)
names = {24: "tma:qk_issue", 26: "tma:v_issue", 25: "mma:kv_full", 0: "mma:S0", 1: "mma:S1", 2: "mma:PA0", 3: "mma:PA1", 4: "mma:PB0", 5: "mma:PB1",
         8: "sm0:s_ready", 9: "sm0:pa", 10: "sm0:pb", 11: "ep0:o_ready", 12: "ep0:released",
         16: "sm1:s_ready", 17: "sm1:pa", 18: "sm1:pb", 19: "ep1:o_ready", 20: "ep1:released",
         6: "sm0q0:c0|sb_seen", 7: "sm0q0:c1", 27: "sm0q0:c2", 28: "sm0q0:c3", 29: "sm0q0:c4", 30: "sm0q0:c5", 31: "sm0q0:c6",
         13: "sm0q0:pa_stwait<", 14: "sm0q0:pa_stwait>", 15: "sm0q0:pb_stwait<", 21: "sm0q0:pb_stwait>", 22: "sm0q0:c2_ldwait<", 23: "sm0q0:c2_ldwait>"}
t0 = int(d[4, 0])
for k in range(4, 10):
    ev = sorted((int(d[k, s]) - t0, n) for s, n in names.items() if int(d[k, s]) != 0)
    print(f"item {k}: " + "  ".join(f"{n}@{v}" for v, n in ev))
