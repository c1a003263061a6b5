import sys, torch
sys.path.insert(0, "/root/repo")
import vimoclip_b200 as vmc
dev = torch.device("cuda:0")
torch.manual_seed(0)
pipe = vmc.ViMoCLIPPipeline("openai/clip-vit-base-patch16", "ViT-B/32", num_classes=140, device=dev, clips_per_step=128)
pipe.rgb.visual.frames_in_flight = 2048
pipe.student.visual_encoder.frames_in_flight = 2048
rgb = torch.randint(0, 256, (256, 16, 3, 224, 224), dtype=torch.uint8).pin_memory()
mot = torch.randint(0, 256, (256, 15, 3, 224, 224), dtype=torch.uint8).pin_memory()
def step():
    lg, er, em = pipe(rgb, mot)
    return lg.cpu(), er.cpu(), em.cpu()
def timed(n=4):
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rnd in range(3):
    for ramp in (False, True):
        pipe.ramp_first_chunk = ramp
        print(f"ramp_first_chunk={ramp}: {timed():.1f} ms per 256-clip e2e step", flush=True)
