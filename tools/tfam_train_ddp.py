"""TFAM training step under DistributedDataParallel (TFAM/train_and_eval.py:66-101, :392).

Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/tfam_train_ddp.py
Each rank owns its own clips (weak scaling); the only communication is DDP's bucketed NCCL all-reduce of the 17.5 M
fp32 gradients, overlapped with the backward kernels by DDP's hooks on the parameters.  Checks, then times:
  1. gradients after a DDP backward == mean over ranks of the local (non-DDP) gradients (dropout off);
  2. 10 AdamW(lr 1e-4, wd 0.1) steps with the reference's dropout rates: finite, decreasing loss;
  3. ms / step and clips / s for `--clips` clips per rank.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def build(dropout, mlp_dropout):
        torch.manual_seed(0)  # same initial weights on every rank
        return vmc.AMO_CLIP(dropout=dropout, mlp_dropout=mlp_dropout, device=dev).to(dev).train()

    gen = torch.Generator().manual_seed(100 + rank)  # different clips per rank
    B = args.clips
    rgb = torch.randn(B, 16, 512, generator=gen).to(dev)
    mot = torch.randn(B, 15, 512, generator=gen).to(dev)
    labels = (torch.rand(B, 140, generator=gen) < 0.05).float().to(dev)
    crit = torch.nn.BCEWithLogitsLoss()

    # ---- 1. DDP gradients == mean of local gradients ----
    local_model = build(0.0, 0.0)
    crit(local_model(rgb[:32], mot[:32]), labels[:32]).backward()
    used = [(n, p) for n, p in local_model.named_parameters() if p.grad is not None]
    expect = torch.cat([p.grad.flatten() for _, p in used])
    dist.all_reduce(expect)
    expect /= world
    ddp = torch.nn.parallel.DistributedDataParallel(build(0.0, 0.0), device_ids=[local], find_unused_parameters=True)
    crit(ddp(rgb[:32], mot[:32]), labels[:32]).backward()
    got = torch.cat([p.grad.flatten() for n, p in ddp.module.named_parameters() if p.grad is not None])
    rel = ((got - expect).norm() / expect.norm()).item()
    assert got.numel() == expect.numel() and rel < 1e-5, rel
    chk = got.double().sum().reshape(1)
    all_chk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    assert all(torch.equal(c, all_chk[0]) for c in all_chk), "ranks disagree after the all-reduce"

    # ---- 2./3. training loop with the reference's optimiser and dropout ----
    ddp = torch.nn.parallel.DistributedDataParallel(build(0.1, 0.3), device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, weight_decay=0.1)
    losses = []
    for step in range(3 + args.steps):
        if step == 3:
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        opt.zero_grad()
        loss = crit(ddp(rgb, mot), labels)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    if rank == 0:
        print(json.dumps({"what": "TFAM training step under DDP (NCCL all-reduce of 17.5 M fp32 gradients)", "n_gpus": world,
                          "clips_per_gpu": B, "ms_per_step": float(ms.item()), "clips_per_s": world * B / (float(ms.item()) / 1e3),
                          "ddp_grad_rel_err_vs_mean_of_local": rel, "loss_first": losses[0], "loss_last": losses[-1],
                          "kernels_launched_per_step": vmc.ops.launch_count() // (3 + args.steps + 2)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
