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
