"""CPU oracle for the CLIP text encoder.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference does not define this network itself: it converts and calls ``transformers.CLIPTextModel``
(``torch2coreml.py:408-441``: ``pipe.text_encoder``; ``pipeline.py:151-175``: ``text_encoder(input_ids=...)
["last_hidden_state"]``).  The oracle therefore IS that library class (the ``transformers`` wheel of this image;
the reference's requirements pin ``transformers==4.44.2``, the image carries a newer release with the same
module), instantiated from a config dict and a state dict, run in fp32 on the CPU.  A plain restatement
(:func:`clip_text_forward`) is kept next to it and checked against the library in ``tests/test_oracle.py`` so
that the parity test still has a checker where ``transformers`` is missing.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def available() -> bool:
    try:
        import transformers  # noqa: F401
        return True
    except Exception:  # pragma: no cover
        return False


def build_clip_text_model(cfg: dict, state_dict: dict):
    from transformers import CLIPTextConfig, CLIPTextModel

    conf = CLIPTextConfig(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"],
                          intermediate_size=cfg["intermediate_size"], num_hidden_layers=cfg["num_hidden_layers"],
                          num_attention_heads=cfg["num_attention_heads"],
                          max_position_embeddings=cfg["max_position_embeddings"], hidden_act=cfg["hidden_act"],
                          layer_norm_eps=cfg.get("layer_norm_eps", 1e-5), projection_dim=cfg["hidden_size"],
                          bos_token_id=cfg["vocab_size"] - 2, eos_token_id=cfg["vocab_size"] - 1, pad_token_id=1)
    conf._attn_implementation = "eager"
    model = CLIPTextModel(conf).eval()
    missing, unexpected = model.load_state_dict({k: v.float() for k, v in state_dict.items()}, strict=False)
    bad = [k for k in missing if "position_ids" not in k]
    if bad or unexpected:
        raise RuntimeError(f"state dict mismatch: missing {bad} unexpected {unexpected}")
    return model


def library_forward(cfg, state_dict, input_ids):
    """input_ids: integer tensor [B, S] -> fp32 last_hidden_state [B, S, D] from transformers.CLIPTextModel."""
    model = build_clip_text_model(cfg, state_dict)
    with torch.no_grad():
        return model(input_ids=input_ids.long())["last_hidden_state"].float()


def clip_text_forward(cfg, sd, input_ids, return_all=False):
    """Restatement of transformers' CLIPTextTransformer.forward (modeling_clip.py): token + position embedding,
    pre-LN blocks with causal self-attention and an MLP, final LayerNorm.  fp32.
    return_all: dict with last_hidden_state, hidden_states (embeddings + one entry per layer, before the final
    LayerNorm), pooler_output (last_hidden_state at the first end-of-text position) and, when the state dict has a
    text_projection, text_embeds -- the outputs torch2coreml.py:416-433 selects from."""
    d, heads, eps = cfg["hidden_size"], cfg["num_attention_heads"], cfg.get("layer_norm_eps", 1e-5)
    ids = input_ids.long()
    b, s = ids.shape
    f = {k: v.float() for k, v in sd.items()}
    x = f["text_model.embeddings.token_embedding.weight"][ids] + f["text_model.embeddings.position_embedding.weight"][:s]
    causal = torch.full((s, s), float("-inf")).triu(1)
    act = (lambda t: t * torch.sigmoid(1.702 * t)) if cfg["hidden_act"] == "quick_gelu" else F.gelu
    hidden_states = [x]
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}."
        h = F.layer_norm(x, (d,), f[p + "layer_norm1.weight"], f[p + "layer_norm1.bias"], eps)
        q, k, v = (F.linear(h, f[p + f"self_attn.{n}.weight"], f[p + f"self_attn.{n}.bias"])
                   .view(b, s, heads, d // heads).transpose(1, 2) for n in ("q_proj", "k_proj", "v_proj"))
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d // heads) + causal, dim=-1) @ v
        x = x + F.linear(att.transpose(1, 2).reshape(b, s, d), f[p + "self_attn.out_proj.weight"], f[p + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (d,), f[p + "layer_norm2.weight"], f[p + "layer_norm2.bias"], eps)
        x = x + F.linear(act(F.linear(h, f[p + "mlp.fc1.weight"], f[p + "mlp.fc1.bias"])), f[p + "mlp.fc2.weight"], f[p + "mlp.fc2.bias"])
        hidden_states.append(x)
    last = F.layer_norm(x, (d,), f["text_model.final_layer_norm.weight"], f["text_model.final_layer_norm.bias"], eps)
    if not return_all:
        return last
    eos = cfg.get("eos_token_id", cfg["vocab_size"] - 1)
    pos = eos_positions(ids, eos)
    out = {"last_hidden_state": last, "hidden_states": hidden_states, "pooler_output": last[torch.arange(b), pos]}
    if "text_projection.weight" in f:
        out["text_embeds"] = F.linear(out["pooler_output"], f["text_projection.weight"])
    return out


def eos_positions(ids, eos_token_id):
    """First end-of-text position per row (modeling_clip.py: (input_ids == eos_token_id).int().argmax(-1));
    rows without one fall back to the position of the largest id (the legacy rule)."""
    ids = ids.long()
    hit = (ids == eos_token_id)
    return torch.where(hit.any(-1), hit.int().argmax(-1), ids.argmax(-1))


def library_forward_all(cfg, state_dict, input_ids):
    """transformers.CLIPTextModel / CLIPTextModelWithProjection with output_hidden_states=True (fp32, CPU)."""
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection

    proj = cfg.get("projection_dim")
    conf = CLIPTextConfig(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"],
                          intermediate_size=cfg["intermediate_size"], num_hidden_layers=cfg["num_hidden_layers"],
                          num_attention_heads=cfg["num_attention_heads"],
                          max_position_embeddings=cfg["max_position_embeddings"], hidden_act=cfg["hidden_act"],
                          layer_norm_eps=cfg.get("layer_norm_eps", 1e-5), projection_dim=proj or cfg["hidden_size"],
                          bos_token_id=cfg["vocab_size"] - 2, eos_token_id=cfg["vocab_size"] - 1, pad_token_id=1)
    conf._attn_implementation = "eager"
    model = (CLIPTextModelWithProjection if proj else CLIPTextModel)(conf).eval()
    missing, unexpected = model.load_state_dict({k: v.float() for k, v in state_dict.items()}, strict=False)
    bad = [k for k in missing if "position_ids" not in k]
    if bad or unexpected:
        raise RuntimeError(f"state dict mismatch: missing {bad} unexpected {unexpected}")
    with torch.no_grad():
        o = model(input_ids=input_ids.long(), output_hidden_states=True)
    out = {"last_hidden_state": o.last_hidden_state.float(), "hidden_states": [h.float() for h in o.hidden_states]}
    if proj:
        out["text_embeds"] = o.text_embeds.float()
    else:
        out["pooler_output"] = o.pooler_output.float()
    return out
