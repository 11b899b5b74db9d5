"""GPU check of the full-size SD-2.1 ControlNet against tests/golden/controlnet_sd21.npz (reference residuals)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from b200sd import config  # noqa: E402
from b200sd.controlnet import ControlNetModel  # noqa: E402

gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "controlnet_sd21.npz"))
cfg = config.SD21_CONTROLNET
sd = config.random_state_dict(config.controlnet_param_shapes(cfg), seed=int(gold["weight_seed"]), dtype=torch.float16)
g = torch.Generator().manual_seed(int(gold["input_seed"]))
x = torch.randn(2, 4, 64, 64, generator=g)
c = torch.randn(2, 1024, 1, 77, generator=g)
cond = torch.rand(2, 3, 512, 512, generator=torch.Generator().manual_seed(int(gold["cond_seed"])))
m = ControlNetModel(cfg, sd, batch=2, height=64, width=64)
out = m(sample=x.half().numpy(), timestep=np.array([501.0, 501.0], np.float16), encoder_hidden_states=c.half().numpy(),
        controlnet_cond=cond.half().numpy())
st = int(gold["stride"])
worst = 0.0
for i in range(13):
    ref = gold[f"residual_{i}"].astype(np.float32)
    err = float(np.abs(out[f"additional_residual_{i}"][:, :, ::st, ::st] - ref).max())
    worst = max(worst, err / max(1.0, float(np.abs(ref).max())))
    print(i, f"max_abs={err:.3e} ref_absmax={np.abs(ref).max():.3f}")
print("WORST_REL", worst)
