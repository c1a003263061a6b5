"""Per-kernel-class time of one student training step (forward + backward) through vmc_profile_begin/end."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc
from vimoclip_b200 import _lib
dev = torch.device("cuda:0")
clips, T = int(os.environ.get("CLIPS", "32")), 16
torch.manual_seed(0)
model = vmc.FlowStudentModel("ViT-B/32", device=dev, num_classes=140).train()
gen = torch.Generator().manual_seed(7)
frames = torch.randint(0, 256, (clips, T, 3, 224, 224), dtype=torch.uint8, generator=gen).to(dev)
teacher = torch.randn(clips, T, 512, generator=gen).to(dev)
labels = (torch.rand(clips, 140, generator=gen) < 0.05).float().to(dev)


def step():
    for p in model.parameters():
        p.grad = None
    emb, dis, logits = model(frames)
    loss = vmc.losses.distillation_loss(dis, teacher, mode="cosine") + vmc.losses.classification_loss(logits, labels)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"step {e0.elapsed_time(e1):.1f} ms")
L = _lib.lib()
L.vmc_profile_begin()
step()
n = 6
ms, fl, by, la = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
_lib.check(L.vmc_profile_end(ms, fl, by, la, n), "vmc_profile_end")
for i, name in enumerate(["prologue", "gemm_tcgen05", "attention_vit", "layernorm(+bwd)", "attention_small(+bwd)", "other"]):
    print(f"{name:24s} {ms[i]:8.2f} ms  {la[i]:5d} launches  {fl[i] / max(ms[i], 1e-9) / 1e9:8.1f} TFLOP/s")
