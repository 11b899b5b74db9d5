"""Checkpoint ingestion (host only): diffusers-layout safetensors round trip, schema validation, sharded index."""
import json
import os

import pytest
import torch

from b200sd import checkpoint, config


def test_safetensors_round_trip_and_validation(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    cfg = config.TINY_UNET
    sd = config.random_state_dict(config.unet_param_shapes(cfg), seed=1, dtype=torch.float16)
    os.makedirs(tmp_path / "unet")
    st.save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "unet" / "diffusion_pytorch_model.safetensors"))
    got = checkpoint.load_component(str(tmp_path), "unet", cfg)
    assert list(got) == list(config.unet_param_shapes(cfg)) and all(torch.equal(got[k], sd[k]) for k in sd)
    # a parameter with the wrong shape / a missing parameter are reported, extras are tolerated
    bad = dict(sd)
    k0 = next(k for k in sd if sd[k].dim() == 4)
    bad[k0] = sd[k0][:, :-1]
    with pytest.raises(ValueError, match="shape mismatch"):
        checkpoint.check_state_dict("unet", cfg, bad)
    with pytest.raises(KeyError, match="missing"):
        checkpoint.check_state_dict("unet", cfg, {k: v for k, v in sd.items() if k != k0})
    extra = dict(sd, **{"encoder.junk.weight": torch.zeros(1)})
    assert len(checkpoint.check_state_dict("unet", cfg, extra)) == len(sd)
    with pytest.raises(KeyError, match="unexpected"):
        checkpoint.check_state_dict("unet", cfg, extra, allow_extra=False)


def test_sharded_index_and_linear_as_conv_shapes(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    cfg = config.TINY_CLIP_TEXT
    sd = config.random_clip_text_state_dict(cfg, seed=2)
    keys = list(sd)
    half = len(keys) // 2
    os.makedirs(tmp_path / "text_encoder")
    wm = {}
    for name, part in (("model-00001-of-00002.safetensors", keys[:half]), ("model-00002-of-00002.safetensors", keys[half:])):
        st.save_file({k: sd[k].contiguous() for k in part}, str(tmp_path / "text_encoder" / name))
        wm.update({k: name for k in part})
    (tmp_path / "text_encoder" / "model.safetensors.index.json").write_text(json.dumps({"weight_map": wm}))
    got = checkpoint.load_component(str(tmp_path), "text_encoder", cfg)
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    # [out, in, 1, 1] (the layout after the reference's load_state_dict pre-hooks, unet.py:121-138) == [out, in]
    ucfg = config.TINY_UNET
    usd = config.random_state_dict(config.unet_param_shapes(ucfg), seed=3)
    k = next(k for k, v in usd.items() if v.dim() == 4 and v.shape[2:] == (1, 1))
    usd[k] = usd[k][:, :, 0, 0]
    assert checkpoint.check_state_dict("unet", ucfg, usd)[k].dim() == 2
