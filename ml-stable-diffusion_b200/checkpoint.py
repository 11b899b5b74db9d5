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
    "vae_encoder": C.vae_encoder_param_shapes,
    "text_encoder": C.clip_text_param_shapes,
}
# AutoencoderKL checkpoints written before diffusers 0.18 name the mid-block attention projections query / key / value /
# proj_attn (diffusers remaps them when it loads the file); the engines use the current names
_VAE_ATTN_RENAMES = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}
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


def remap_legacy_vae_keys(sd: dict) -> dict:
    """Deprecated AutoencoderKL attention names -> current ones (values untouched; [C, C, 1, 1] vs [C, C] is accepted
    by the shape check below)."""
    out = {}
    for k, v in sd.items():
        if ".attentions." in k:
            for old, new in _VAE_ATTN_RENAMES.items():
                if old in k:
                    k = k.replace(old, new)
                    break
        out[k] = v
    return out


def check_state_dict(component: str, cfg: dict, sd: dict, allow_extra=True) -> dict:
    """Validates names and shapes against the architecture schema; returns the subset the engine consumes.
    ``vae_decoder`` / ``vae_encoder`` accept a full AutoencoderKL state dict (the other half is dropped)."""
    want = _SCHEMAS[component](cfg)
    if component.startswith("vae"):
        sd = remap_legacy_vae_keys(sd)
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
    sub = {"unet": "unet", "controlnet": "", "vae_decoder": "vae", "vae_encoder": "vae", "text_encoder": "text_encoder",
           "text_encoder_2": "text_encoder_2", "unet_refiner": "unet"}.get(component, component)
    schema = "text_encoder" if component == "text_encoder_2" else ("unet" if component == "unet_refiner" else component)
    return check_state_dict(schema, cfg, read_state_dict(os.path.join(model_dir, sub) if sub else model_dir))


def read_config(model_dir: str, component: str) -> dict:
    """``<model_dir>/<component>/config.json`` (scheduler: ``scheduler_config.json``) of a diffusers pipeline directory,
    reduced to the keys the engines read (private ``_class_name`` / ``_diffusers_version`` entries dropped)."""
    name = "scheduler_config.json" if component == "scheduler" else "config.json"
    with open(os.path.join(model_dir, component, name)) as f:
        cfg = json.load(f)
    return {k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items() if not k.startswith("_")}
