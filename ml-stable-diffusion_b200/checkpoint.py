"""Weight ingestion from diffusers-layout checkpoints (host side).

The reference builds its networks from the diffusers pipeline's modules and loads their ``state_dict()`` unchanged
(``torch2coreml.py:915-918``: ``reference_unet.load_state_dict(pipe.unet.state_dict())``; the pre-hooks of
``unet.py:121-138`` only reshape ``nn.Linear`` weights to 1x1 convolutions).  The engines of this package take those
same parameter names, so ingestion is: read ``<model dir>/<component>/diffusion_pytorch_model.safetensors`` (or a
sharded / ``.bin`` variant), check it against the architecture schema, hand it to the engine.
"""
from __future__ import annotations

import json
import os

import torch

from . import config as C

_SCHEMAS = {
    "unet": C.unet_param_shapes,
    "controlnet": C.controlnet_param_shapes,
    "vae_decoder": C.vae_decoder_param_shapes,
    "text_encoder": C.clip_text_param_shapes,
}
_FILES = ("diffusion_pytorch_model.safetensors", "model.safetensors", "diffusion_pytorch_model.fp16.safetensors",
          "model.fp16.safetensors", "diffusion_pytorch_model.bin", "pytorch_model.bin")


def read_state_dict(path: str) -> dict:
    """A .safetensors / .bin file, a sharded ``*.index.json``, or a component directory holding one of them."""
    if os.path.isdir(path):
        for f in _FILES:
            if os.path.exists(os.path.join(path, f)):
                return read_state_dict(os.path.join(path, f))
        idx = [f for f in os.listdir(path) if f.endswith(".index.json")]
        if idx:
            return read_state_dict(os.path.join(path, idx[0]))
        raise FileNotFoundError(f"no checkpoint file found under {path}")
    if path.endswith(".index.json"):
        with open(path) as f:
            shards = sorted(set(json.load(f)["weight_map"].values()))
        sd = {}
        for s in shards:
            sd.update(read_state_dict(os.path.join(os.path.dirname(path), s)))
        return sd
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu", weights_only=True)


def _canonical(shape):
    """Linear weights may be stored as [out, in] (diffusers) or [out, in, 1, 1] (after the reference's pre-hooks)."""
    shape = tuple(shape)
    return shape[:2] if len(shape) == 4 and shape[2:] == (1, 1) else shape


def check_state_dict(component: str, cfg: dict, sd: dict, allow_extra=True) -> dict:
    """Validates names and shapes against the architecture schema; returns the subset the engine consumes.
    ``vae_decoder`` accepts a full AutoencoderKL state dict (encoder / quant_conv entries are dropped)."""
    want = _SCHEMAS[component](cfg)
    missing = [k for k in want if k not in sd]
    if missing:
        raise KeyError(f"{component}: {len(missing)} parameters missing from the checkpoint, e.g. {missing[:3]}")
    bad = [(k, tuple(sd[k].shape), tuple(want[k])) for k in want if _canonical(sd[k].shape) != _canonical(want[k])]
    if bad:
        raise ValueError(f"{component}: shape mismatch for {len(bad)} parameters, e.g. {bad[:3]}")
    extra = [k for k in sd if k not in want]
    if extra and not allow_extra:
        raise KeyError(f"{component}: unexpected parameters, e.g. {extra[:3]}")
    return {k: sd[k] for k in want}


def load_component(model_dir: str, component: str, cfg: dict) -> dict:
    """``model_dir`` is a diffusers pipeline directory (``unet/``, ``vae/``, ``text_encoder/``, ...)."""
    sub = {"unet": "unet", "controlnet": "", "vae_decoder": "vae", "text_encoder": "text_encoder"}[component]
    return check_state_dict(component, cfg, read_state_dict(os.path.join(model_dir, sub) if sub else model_dir))
