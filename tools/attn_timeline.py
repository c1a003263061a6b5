"""Needs the library built with the stamps: make -C vimo-clip_b200/csrc clean all EXTRA_NVCCFLAGS=-DVMC_ATTN_TIMELINE."""
"""Phase timeline (clock64) of CTA 0 of attention_vit3_kernel; prints cycles relative to the first stamp."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc
from vimoclip_b200 import ops
dev = torch.device("cuda:0")
F_, L, heads = 512, 197, 12
qkv = torch.randn(F_ * L, 3 * heads * 64, device=dev).to(torch.bfloat16)
dbg = torch.zeros(16 * 32, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.attention_vit(qkv, F_, L, heads, impl=3)
ops.set_option(7, dbg.data_ptr())
ops.attention_vit(qkv, F_, L, heads, impl=3)
torch.cuda.synchronize()
ops.set_option(7, 0)
d = dbg.cpu().view(16, 32)
t0 = int(d[0, 0])
names = {0: "mma:item", 1: "mma:kv_full", 2: "mma:S0go", 3: "mma:S1go", 4: "mma:PA0go", 5: "mma:PA1go", 6: "mma:PB0go", 7: "mma:PB1go",
         8: "A:wait_s", 9: "A:s_ready", 10: "A:p_done", 11: "A:o_ready", 12: "A:released",
         16: "B:wait_s", 17: "B:s_ready", 18: "B:p_done", 19: "B:o_ready", 20: "B:released"}
for k in range(2, 8):
    row = {names[s]: int(d[k, s]) - t0 for s in names if int(d[k, s]) != 0}
    base = row.get("mma:item", 0)
    print(f"item {k}: start {base}  " + "  ".join(f"{n}={v - base}" for n, v in row.items() if n != "mma:item"))
