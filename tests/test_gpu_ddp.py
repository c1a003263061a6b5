"""Multi-GPU: the TFAM training step under DistributedDataParallel over NCCL (needs >= 2 GPUs; skipped otherwise)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tfam_training_step_ddp_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "tfam_train_ddp.py"), "--clips", "64", "--steps", "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["ddp_grad_rel_err_vs_mean_of_local"] < 1e-5 and line["loss_last"] < line["loss_first"]


@pytest.mark.gpu
def test_dataparallel_wrapping_matches_direct_call():
    """The reference wraps every model in ``torch.nn.DataParallel`` (train.py:64, inference.py:80,
    TFAM/train_and_eval.py:392) and saves / loads checkpoints with the ``module.`` prefix (train.py:167, inference.py:86)."""
    import vimoclip_b200 as vmc
    from oracle import tfam as otfam, weights

    dev = torch.device("cuda:0")
    o_tf = otfam.TfamOracle().eval()
    with torch.no_grad():
        weights.randomise_tfam_(o_tf, 0)
    tfam = vmc.AMO_CLIP(device=dev)
    tfam.load_state_dict(o_tf.state_dict(), strict=True)
    tfam = tfam.to(dev).eval()
    gen = torch.Generator().manual_seed(2)
    rgb, mot = torch.randn(8, 16, 512, generator=gen).to(dev), torch.randn(8, 15, 512, generator=gen).to(dev)
    direct = tfam(rgb, mot)
    dp = torch.nn.DataParallel(tfam)  # all visible GPUs: scatter along dim 0, replicate, gather on cuda:0
    out = dp(rgb, mot)
    assert out.device == direct.device and (out - direct).abs().max().item() <= 1e-5
    sd = dp.state_dict()
    assert all(k.startswith("module.") for k in sd)
    fresh = torch.nn.DataParallel(vmc.AMO_CLIP(device=dev).to(dev).eval())
    fresh.load_state_dict(sd, strict=True)
    assert (fresh(rgb, mot) - direct).abs().max().item() <= 1e-5

    student = vmc.FrameDiffStudentModel("ViT-B/32", device=dev).eval()
    frames = torch.randint(0, 256, (2, 3, 3, 224, 224), dtype=torch.uint8, generator=gen).to(dev)
    e0, d0, l0 = student(frames)
    e1, d1, l1 = torch.nn.DataParallel(student)(frames)
    cos = torch.nn.functional.cosine_similarity(e0.flatten(0, 1).double(), e1.flatten(0, 1).double(), dim=-1).min().item()
    assert cos >= 0.99999 and (l0 - l1).abs().max().item() <= 1e-3


@pytest.mark.gpu
def test_dataparallel_replicas_reuse_packed_weights_and_train():
    """What torch.nn.DataParallel does on every forward, reproduced on ONE GPU with torch.nn.parallel.replicate: the replicas'
    `_parameters` are empty (their tensors are broadcast copies listed in `_former_parameters`).  (i) eval: a replica computes
    the same outputs and REUSES the packed weights of the wrapped module (no re-pack per forward, inference.py:129 calls the
    model once per 15-frame clip); (ii) train (train.py:64, TFAM/train_and_eval.py:392 wrap the trained model): the backward
    through a replica reaches the real parameters with the same gradients as a direct call."""
    import vimoclip_b200 as vmc
    from torch.nn.parallel import replicate

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    torch.manual_seed(5)
    # ---- TFAM ----
    tfam = vmc.AMO_CLIP(device=dev, dropout=0.0, mlp_dropout=0.0).to(dev).eval()
    rgb, mot = torch.randn(4, 16, 512, generator=gen).to(dev), torch.randn(4, 15, 512, generator=gen).to(dev)
    direct = tfam(rgb, mot)
    packed = tfam._cache.peek(("fused", 0))
    assert packed is not None
    rep = replicate(tfam, [0])[0]
    assert len(list(rep.parameters())) == 0  # the situation the advisor described
    assert torch.equal(rep(rgb, mot), direct)
    assert tfam._cache.peek(("fused", 0)) is packed  # same pack object: nothing was rebuilt
    tfam.train()
    labels = (torch.rand(4, 140, generator=gen) < 0.1).float().to(dev)
    torch.nn.functional.binary_cross_entropy_with_logits(tfam(rgb, mot), labels).backward()
    want = {n: p.grad.clone() for n, p in tfam.named_parameters() if p.grad is not None}
    tfam.zero_grad()
    rep = replicate(tfam, [0])[0]
    torch.nn.functional.binary_cross_entropy_with_logits(rep(rgb, mot), labels).backward()
    got = {n: p.grad for n, p in tfam.named_parameters() if p.grad is not None}
    assert set(want) <= set(got) and len(want) > 50
    for n in set(got) - set(want):  # Broadcast.backward hands the parameters this mode does not use a zero gradient
        assert float(got[n].abs().max()) == 0.0, n
    for n in want:
        assert torch.allclose(got[n], want[n], rtol=1e-4, atol=1e-6), n
    # ---- student ----
    student = vmc.FrameDiffStudentModel("ViT-B/32", device=dev).eval()
    frames = torch.randint(0, 256, (2, 2, 3, 224, 224), dtype=torch.uint8, generator=gen).to(dev)
    e0, d0, l0 = student(frames)
    pack = student.visual_encoder._cache.peek(("pack", 0))
    rep = replicate(student, [0])[0]
    e1, d1, l1 = rep(frames)
    assert torch.equal(e0, e1) and torch.equal(l0, l1)
    assert student.visual_encoder._cache.peek(("pack", 0)) is pack
    student.train()
    rep = replicate(student, [0])[0]
    e, d, lg = rep(frames)
    (d.square().mean() + lg.square().mean()).backward()
    grads = [p.grad for p in student.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads) and len(grads) == 160


@pytest.mark.gpu
def test_dataparallel_two_gpus_eval_and_train():
    """A real 2-GPU torch.nn.DataParallel run (the reference's wrapping): eval outputs equal the direct call, a training
    step produces gradients on the wrapped module, and the pipeline's sharded forward gathers in clip order."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import vimoclip_b200 as vmc

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(6)
    torch.manual_seed(6)
    tfam = vmc.AMO_CLIP(device=dev, dropout=0.0, mlp_dropout=0.0).to(dev).eval()
    rgb, mot = torch.randn(8, 16, 512, generator=gen).to(dev), torch.randn(8, 15, 512, generator=gen).to(dev)
    direct = tfam(rgb, mot)
    dp = torch.nn.DataParallel(tfam, device_ids=[0, 1])
    for _ in range(2):  # second call: packs cached per device
        out = dp(rgb, mot)
        assert (out - direct).abs().max().item() <= 1e-5
    assert tfam._cache.peek(("fused", 1)) is not None
    dp.train()
    labels = (torch.rand(8, 140, generator=gen) < 0.1).float().to(dev)
    torch.nn.functional.binary_cross_entropy_with_logits(dp(rgb, mot), labels).backward()
    assert all(p.grad is not None for n, p in tfam.named_parameters() if "projection_layer" not in n)
    student = vmc.FrameDiffStudentModel("ViT-B/32", device=dev).eval()
    frames = torch.randint(0, 256, (4, 3, 3, 224, 224), dtype=torch.uint8, generator=gen).to(dev)
    e0, _, l0 = student(frames)
    e1, _, l1 = torch.nn.DataParallel(student, device_ids=[0, 1])(frames)
    cos = torch.nn.functional.cosine_similarity(e0.flatten(0, 1).double(), e1.flatten(0, 1).double(), dim=-1).min().item()
    assert cos >= 0.99999 and (l0 - l1).abs().max().item() <= 1e-3


@pytest.mark.gpu
def test_forward_sharded_two_ranks():
    """ViMoCLIPPipeline.forward_sharded under a 2-rank NCCL spawn: rank r encodes clips r::2, every rank ends with the full,
    clip-ordered logits / embeddings, equal to the single-GPU forward."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tools", "sharded_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["world"] == 2 and line["max_abs_logits"] <= 1e-5 and line["max_abs_emb"] <= 1e-5
