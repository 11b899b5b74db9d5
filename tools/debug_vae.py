import os, sys
os.environ["B200SD_DEBUG_SYNC"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200sd import config, lib
from b200sd.vae import VAEDecoderModel
from oracle import restated as R
vcfg = config.TINY_VAE
vsd = config.random_state_dict(config.vae_decoder_param_shapes(vcfg), seed=3)
vae = VAEDecoderModel(vcfg, vsd, batch=1, height=16, width=16, device="cuda:0")
z = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(1))
print("calling vae", flush=True)
img = vae(z=z.half().numpy())["image"]
with torch.no_grad():
    iref = R.vae_decode({k_: v.half().float() for k_, v in vsd.items()}, vcfg, z.half().float()).numpy()
print("vae max err", float(np.abs(img - iref).max()), "ref absmax", float(np.abs(iref).max()), flush=True)
