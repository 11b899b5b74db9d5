"""GPU tests of the checkpoint -> pipeline factory (the ``get_coreml_pipe`` counterpart, pipeline.py:607-697) and the SDXL
refiner hand-off inside the device loop (StableDiffusionXLPipeline.swift:205-225), on tiny random-init models."""
import json
import os

import numpy as np
import pytest
import torch

from b200sd import config
from oracle import restated as R

pytestmark = pytest.mark.gpu


def _write_component(root, name, sd, cfg, cls=None):
    st = pytest.importorskip("safetensors.torch")
    os.makedirs(root / name, exist_ok=True)
    st.save_file({k: v.contiguous() for k, v in sd.items()}, str(root / name / "diffusion_pytorch_model.safetensors"))
    meta = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    if cls:
        meta["_class_name"] = cls
    (root / name / "config.json").write_text(json.dumps(meta))


def _model_dir(tmp_path, ucfg, seed):
    usd = config.random_state_dict(config.unet_param_shapes(ucfg), seed=seed, dtype=torch.float16)
    vcfg = config.TINY_VAE
    vsd = config.random_state_dict(config.vae_decoder_param_shapes(vcfg), seed=seed + 1, dtype=torch.float16)
    # legacy AutoencoderKL attention names (pre-0.18 checkpoints): the reader must remap them
    legacy = {}
    for k, v in vsd.items():
        for new, old in ((".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn.")):
            if ".attentions." in k and new in k:
                k = k.replace(new, old)
        legacy[k] = v
    _write_component(tmp_path, "unet", usd, ucfg, "UNet2DConditionModel")
    _write_component(tmp_path, "vae", legacy, vcfg, "AutoencoderKL")
    os.makedirs(tmp_path / "scheduler", exist_ok=True)
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps({"_class_name": "DDIMScheduler"}))
    return usd, vsd


def test_from_pretrained_equals_direct_construction(cuda_lib, tmp_path):
    from b200sd.model import UNetModel
    from b200sd.pipeline import B200StableDiffusionPipeline
    from b200sd.vae import VAEDecoderModel

    ucfg = config.TINY_UNET
    usd, vsd = _model_dir(tmp_path, ucfg, seed=11)
    pipe = B200StableDiffusionPipeline.from_pretrained(str(tmp_path), height=64, width=64)
    assert pipe.scheduler_name == "DDIM" and not pipe.xl and pipe.force_zeros_for_empty_prompt is False
    direct = B200StableDiffusionPipeline(UNetModel(ucfg, usd, batch=2, height=16, width=16),
                                         VAEDecoderModel(config.TINY_VAE, vsd, batch=1, height=16, width=16),
                                         scheduler="DDIM", force_zeros_for_empty_prompt=False)
    kw = dict(height=64, width=64, num_inference_steps=3, guidance_scale=5.0, output_type="np", seed=7, rng="torch")
    a = pipe("a red cube", **kw).images
    b = direct("a red cube", **kw).images
    assert a.shape == (1, 64, 64, 3) and np.array_equal(a, b)
    with pytest.raises(ValueError, match="not implemented"):
        (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps({"_class_name": "EulerDiscreteScheduler"}))
        B200StableDiffusionPipeline.from_pretrained(str(tmp_path), height=64, width=64)


def test_sdxl_refiner_hand_off_vs_oracle_loop(cuda_lib):
    """Base UNet for the first int(n * refiner_start) steps, then the refiner with its own conditioning: hidden states,
    pooled states, and (original size, crop, aesthetic score) geometry, negative score on the unconditional row."""
    from b200sd.model import UNetModel
    from b200sd.pipeline import B200StableDiffusionPipeline
    from b200sd.vae import VAEDecoderModel

    bcfg = config.TINY_XL_UNET
    rcfg = dict(bcfg, projection_class_embeddings_input_dim=64 + 5 * 32, num_time_ids=5)
    bsd = config.random_state_dict(config.unet_param_shapes(bcfg), seed=21, dtype=torch.float16)
    rsd = config.random_state_dict(config.unet_param_shapes(rcfg), seed=22, dtype=torch.float16)
    vsd = config.random_state_dict(config.vae_decoder_param_shapes(config.TINY_VAE), seed=23, dtype=torch.float16)
    pipe = B200StableDiffusionPipeline(UNetModel(bcfg, bsd, batch=2, height=16, width=16),
                                       VAEDecoderModel(config.TINY_VAE, vsd, batch=1, height=16, width=16),
                                       scheduler="DDIM", xl=True, unet_refiner=UNetModel(rcfg, rsd, batch=2, height=16, width=16))
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(2, 96, 1, 77, generator=g).half()
    pooled = torch.randn(2, 64, generator=g)
    remb = torch.randn(2, 96, 1, 77, generator=g).half()
    rpooled = torch.randn(2, 64, generator=g)
    lat0 = torch.randn(1, 4, 16, 16, generator=g).half().float()
    steps, gs, rstart = 5, 4.0, 0.6
    tid = torch.tensor([[64.0, 64.0, 0.0, 0.0, 64.0, 64.0]] * 2)
    rtid = torch.tensor([[64.0, 64.0, 0.0, 0.0, 2.5], [64.0, 64.0, 0.0, 0.0, 6.0]])
    final = pipe.denoise(emb, lat0, steps, gs, time_ids=tid, text_embeds=pooled,
                         refiner={"encoder_hidden_states": remb, "time_ids": rtid, "text_embeds": rpooled},
                         refiner_start=rstart).cpu().clone()
    # oracle loop
    abar = R.alphas_cumprod()
    x = lat0.clone()
    switch = int(np.float32(steps) * np.float32(rstart))
    assert switch == 3
    with torch.no_grad():
        for i, t in enumerate(R.leading_timesteps(steps)):
            tt = torch.tensor([float(t)] * 2)
            xin = torch.cat([x, x]).half().float()
            if i < switch:
                eps = R.unet_forward(bsd, bcfg, xin, tt, emb, time_ids=tid, text_embeds=pooled)
            else:
                eps = R.unet_forward(rsd, rcfg, xin, tt, remb, time_ids=rtid, text_embeds=rpooled)
            x = R.ddim_step(R.cfg_combine(eps[:1], eps[1:], gs), t, x, abar, steps)
    rel = float((final - x).abs().max() / x.abs().max())
    print(f"refiner hand-off: latent rel err after {steps} steps = {rel:.3e}")
    assert rel < 3e-2
    # without refiner inputs the base UNet runs every step: a different result
    plain = pipe.denoise(emb, lat0, steps, gs, time_ids=tid, text_embeds=pooled).cpu()
    assert not torch.allclose(plain, final)
