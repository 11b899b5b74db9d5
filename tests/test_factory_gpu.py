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


def test_sdxl_call_through_both_text_encoders(cuda_lib, tmp_path):
    """SDXL ``__call__`` from a model directory with two tokenizers and two text encoders (pipeline.py:123-257): the
    prompt goes through both BPE tokenizers and both encoders on the device, ``hidden_embeds`` are concatenated on the
    feature axis (encoder 1 first), the pooled / projected output of encoder 2 conditions ``text_time``, negatives are
    zeros.  Checked against the same call fed with embeddings computed by the oracle's CLIP restatement."""
    from b200sd.pipeline import B200StableDiffusionPipeline
    from oracle import clip_text

    st = pytest.importorskip("safetensors.torch")
    c1, c2 = config.TINY_CLIP_TEXT, config.TINY_CLIP_TEXT_PROJ
    ucfg = dict(config.TINY_XL_UNET, cross_attention_dim=c1["hidden_size"] + c2["hidden_size"],
                projection_class_embeddings_input_dim=c2["projection_dim"] + 6 * 32)
    _model_dir(tmp_path, ucfg, seed=31)
    sds = {}
    for name, cfg, seed in (("text_encoder", c1, 41), ("text_encoder_2", c2, 42)):
        sds[name] = config.random_clip_text_state_dict(cfg, seed=seed, dtype=torch.float16)
        os.makedirs(tmp_path / name, exist_ok=True)
        st.save_file({k: v.contiguous() for k, v in sds[name].items()}, str(tmp_path / name / "model.safetensors"))
        (tmp_path / name / "config.json").write_text(json.dumps(dict(cfg, _class_name="CLIPTextModel")))
    vocab = {"<|startoftext|>": 998, "<|endoftext|>": 999, "low</w>": 2, "er</w>": 3, "new": 4, "lo": 5, "w": 6, "!": 0}
    for tdir, extra in (("tokenizer", {}), ("tokenizer_2", {"pad_token": "!"})):
        os.makedirs(tmp_path / tdir, exist_ok=True)
        (tmp_path / tdir / "merges.txt").write_text("#version: 0.2\nl o\nlo w</w>\ne r</w>\nn e\nne w\n")
        (tmp_path / tdir / "vocab.json").write_text(json.dumps(vocab))
        if extra:
            (tmp_path / tdir / "special_tokens_map.json").write_text(json.dumps({"pad_token": {"content": extra["pad_token"]}}))
    pipe = B200StableDiffusionPipeline.from_pretrained(str(tmp_path), height=64, width=64)
    assert pipe.xl and pipe.force_zeros_for_empty_prompt and pipe.text_encoder_2 is not None and pipe.tokenizer_2 is not None
    assert pipe.tokenizer_2.pad_token == "!" and pipe.tokenizer.pad_token == "<|endoftext|>"
    assert list(np.asarray(pipe.tokenizer_2("low"))[0][:5]) == [998, 2, 999, 0, 0]
    kw = dict(height=64, width=64, num_inference_steps=3, guidance_scale=5.0, output_type="np", seed=9, rng="torch")
    img = pipe("low newer", **kw).images
    assert img.shape == (1, 64, 64, 3) and np.isfinite(img).all()
    # the same call with the oracle's embeddings for the same token ids
    hid, pooled = [], None
    for name, cfg, tok in (("text_encoder", c1, pipe.tokenizer), ("text_encoder_2", c2, pipe.tokenizer_2)):
        ids = torch.from_numpy(np.asarray(tok("low newer"))).long()
        with torch.no_grad():
            ref = clip_text.clip_text_forward(cfg, sds[name], ids, return_all=True)
        hid.append(ref["hidden_states"][-2])
        pooled = ref["text_embeds" if cfg.get("projection_dim") else "pooler_output"]
    emb = torch.cat(hid, -1).permute(0, 2, 1)[:, :, None, :]                       # [1, 256, 1, 77]
    emb = torch.cat([torch.zeros_like(emb), emb]).half().numpy()                   # zero negatives, uncond half first
    pool = torch.cat([torch.zeros_like(pooled), pooled]).float().numpy()
    ref_img = pipe("low newer", prompt_embeds=emb, pooled_prompt_embeds=pool, **kw).images
    err = float(np.abs(img.astype(np.float32) - ref_img.astype(np.float32)).max())
    print(f"SDXL __call__ through both encoders vs oracle embeddings: max image diff {err:.4f}")
    assert err <= 0.03
