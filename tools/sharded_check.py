"""2+ rank check of ViMoCLIPPipeline.forward_sharded (torchrun): sharded + gathered == single-GPU forward on the same clips."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimoclip_b200 as vmc  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)  # same weights on every rank
pipe = vmc.ViMoCLIPPipeline("ViT-B/16", "ViT-B/32", device=dev, clips_per_step=4)
gen = torch.Generator().manual_seed(3)
N = 7  # ragged: 4 + 3 clips
rgb = torch.randint(0, 256, (N, 4, 3, 224, 224), dtype=torch.uint8, generator=gen)
mot = torch.randint(0, 256, (N, 3, 3, 224, 224), dtype=torch.uint8, generator=gen)
ids = torch.from_numpy(vmc.indexing.shard_ids(N, rank, world))
lg, er, em = pipe.forward_sharded(rgb[ids].to(dev), mot[ids].to(dev), N)
lg0, er0, em0 = pipe(rgb.to(dev), mot.to(dev))
res = torch.tensor([(lg - lg0).abs().max(), max((er - er0).abs().max(), (em - em0).abs().max())], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "max_abs_logits": float(res[0]), "max_abs_emb": float(res[1])}))
dist.destroy_process_group()
