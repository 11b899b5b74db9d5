"""Extra data points for BASELINE configs[3] (SDXL-base 768x768) and configs[4] (SD-2.1 + ControlNet): UNet
iterations per second at batch 2 with random-init weights, CUDA-graph replay, CUDA events."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from b200sd import config, lib as L  # noqa: E402
from b200sd.controlnet import ControlNetModel  # noqa: E402
from b200sd.model import UNetModel  # noqa: E402


def rand_sd(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: ((torch.rand(v, generator=g) * 2 - 1) * (v[1] * v[2] * v[3]) ** -0.5).half() if len(v) == 4 else
            (torch.ones(v) if k.endswith("weight") else torch.zeros(v)).half() for k, v in shapes.items()}


def time_it(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {}
# ---- config 4: SDXL-base, 768x768 -> 96x96 latents ----
cfg = config.SDXL_BASE_UNET
m = UNetModel(cfg, rand_sd(config.unet_param_shapes(cfg), 1), batch=2, height=96, width=96)
x = torch.randn(2, 4, 96, 96, device="cuda")
t = torch.full((2,), 981.0, device="cuda")
c = torch.randn(2, 2048, 1, 77, device="cuda").half()
tid = torch.tensor([[768.0, 768, 0, 0, 768, 768]] * 2, device="cuda")
te = torch.randn(2, 1280, device="cuda")
ms = time_it(lambda: m.forward_device(x, t, c, tid, te))
assert torch.isfinite(m._out).all()
out["sdxl_768"] = {"ms_per_iter": round(ms, 3), "iter_per_s": round(1e3 / ms, 2), "tflops": round(7.282 / ms * 1e3, 1),
                   "launches": m.launches_per_call}
del m
torch.cuda.empty_cache()
# ---- config 5: SD-2.1 + ControlNet ----
ucfg = dict(config.SD21_BASE_UNET, support_controlnet=True)
unet = UNetModel(ucfg, rand_sd(config.unet_param_shapes(ucfg), 2), batch=2, height=64, width=64)
cn = ControlNetModel(config.SD21_CONTROLNET, rand_sd(config.controlnet_param_shapes(config.SD21_CONTROLNET), 3),
                     batch=2, height=64, width=64)
x = torch.randn(2, 4, 64, 64, device="cuda")
c = torch.randn(2, 1024, 1, 77, device="cuda").half()
cn._sample.copy_(x)
cn._t.fill_(981.0)
cn._ctx.copy_(c)
cn._cond.uniform_()


def step():
    res = cn.forward_device()                       # NHWC fp16 residuals stay on the device
    nchw = [r.permute(0, 3, 1, 2) for r in res]     # views; UNetModel copies them into its static buffers
    unet.forward_device(x, t, c, additional_residuals=nchw)


ms = time_it(step)
assert torch.isfinite(unet._out).all()
out["sd21_controlnet"] = {"ms_per_iter": round(ms, 3), "iter_per_s": round(1e3 / ms, 2),
                          "tflops": round((1.609 + 0.567) / ms * 1e3, 1),
                          "note": "ControlNet and UNet both CUDA-graph replays"}
print(json.dumps(out))
