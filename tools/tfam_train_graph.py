"""TFAM training step (TFAM/train_and_eval.py:66-101: forward, BCEWithLogits, backward, AdamW) as ONE CUDA graph vs eager launches.
The step is ~400 kernel launches of a few microseconds each; forward + loss + backward + optimiser are captured once
(torch.cuda.graph; our kernels are enqueued on the capturing stream, AdamW with capturable=True) and replayed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("CLIPS", "256"))
torch.manual_seed(0)
gen = torch.Generator().manual_seed(1)


def make():
    torch.manual_seed(0)
    model = vmc.AMO_CLIP(num_classes=140, device=dev).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.1, capturable=True)
    return model, opt


rgb = torch.randn(B, 16, 512, generator=gen).to(dev)
mot = torch.randn(B, 15, 512, generator=gen).to(dev)
lens_r = torch.randint(8, 17, (B,), generator=gen)
lens_m = (lens_r - 1).clamp(min=1)
m_r = (torch.arange(16)[None, :] < lens_r[:, None]).to(dev)
m_m = (torch.arange(15)[None, :] < lens_m[:, None]).to(dev)
labels = (torch.rand(B, 140, generator=gen) < 0.03).float().to(dev)
crit = torch.nn.BCEWithLogitsLoss()


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


model, opt = make()


def eager_step():
    opt.zero_grad(set_to_none=True)
    loss = crit(model(rgb, mot, m_r, m_m), labels)
    loss.backward()
    opt.step()
    return loss


t_eager = timeit(eager_step)
l_eager = [float(eager_step()) for _ in range(3)]

model, opt = make()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        crit(model(rgb, mot, m_r, m_m), labels).backward()
        opt.step()
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
opt.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    static_loss = crit(model(rgb, mot, m_r, m_m), labels)
    static_loss.backward()
    opt.step()
t_graph = timeit(g.replay)
losses = []
for _ in range(40):
    g.replay()
    losses.append(float(static_loss))
print(f"TFAM training step, {B} clips: eager launches {t_eager:.2f} ms, one CUDA graph {t_graph:.2f} ms; "
      f"loss under replay {losses[0]:.4f} -> {losses[-1]:.4f} (eager after its timing loop: {l_eager[-1]:.4f})")
