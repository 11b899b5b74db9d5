"""Per-class step time of the batched configuration (BASELINE configs[2] shape: 8 prompts per GPU, UNet batch 16) with
bench.py's own loop graphs: is any kernel class badly tuned at the larger batch?"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import b200sd  # noqa: E402,F401
from b200sd.pipeline import B200StableDiffusionPipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
pipe = B200StableDiffusionPipeline.from_random_init("sd21-base", images_per_call=n, device=dev, seed=1, scheduler="DDIM")
g = torch.Generator().manual_seed(3)
emb = torch.cat([torch.zeros(n, 1024, 1, 77), torch.randn(n, 1024, 1, 77, generator=g)]).half()
pipe.unet._ctx.copy_(emb)
lat0 = torch.randn(n, 4, 64, 64, generator=g).half().float().to(dev)
loop = bench.LoopBench(pipe, lat0)


def barrier():
    torch.cuda.synchronize()


full = bench.timed_replays(loop.capture(bench.N_STEPS_IMG), 3, barrier) / bench.N_STEPS_IMG
out = {"images_per_call": n, "ms_per_step": round(full, 3), "launches_per_step": round(loop.launches_per_image / bench.N_STEPS_IMG, 1)}
for name, bit in (("gemm_conv", 1), ("attention", 2), ("normalisation", 4), ("elementwise", 8)):
    out[name] = round(bench.timed_replays(loop.capture(bench.N_STEPS_IMG, classes=bit), 3, barrier) / bench.N_STEPS_IMG, 3)
out["unet_tflop_per_s"] = round(1.609 * n / (full * 1e-3), 1)
print(json.dumps(out), flush=True)
