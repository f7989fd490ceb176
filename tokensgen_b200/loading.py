"""Checkpoint loading for the mirrors: the diffusers directory layout the reference loads with `from_pretrained`
(infer_cogvideo_mp_fifo.py:150-176): `<root>/<subfolder>/config.json` + `diffusion_pytorch_model*.safetensors` (single
file or sharded with an index json) or `.bin`/`.pt`.  Tensors go host -> device once, in the requested dtype; nothing is
pickled between processes."""
from __future__ import annotations

import inspect
import json
import os
from typing import Any, Dict, Optional

import torch


def load_config(root: str, subfolder: Optional[str] = None, name: str = "config.json") -> Dict[str, Any]:
    path = os.path.join(root, subfolder or "", name)
    with open(path) as f:
        cfg = json.load(f)
    return {k: v for k, v in cfg.items() if not k.startswith("_")}


def load_state_dict(root: str, subfolder: Optional[str] = None, dtype: Optional[torch.dtype] = None) -> Dict[str, torch.Tensor]:
    d = os.path.join(root, subfolder or "")
    files = sorted(os.listdir(d))
    sd: Dict[str, torch.Tensor] = {}
    index = [f for f in files if f.endswith(".safetensors.index.json")]
    if index:
        with open(os.path.join(d, index[0])) as f:
            shards = sorted(set(json.load(f)["weight_map"].values()))
    else:
        shards = [f for f in files if f.endswith(".safetensors")]
    if shards:
        from safetensors.torch import load_file
        for s in shards:
            sd.update(load_file(os.path.join(d, s)))
    else:
        blobs = [f for f in files if f.endswith((".bin", ".pt", ".pth")) and "optimizer" not in f]
        if not blobs:
            raise FileNotFoundError(f"no weights (*.safetensors / *.bin / *.pt) under {d}")
        for b in blobs:
            sd.update(torch.load(os.path.join(d, b), map_location="cpu", weights_only=True))
    if dtype is not None:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    return sd


def build_from_pretrained(cls, root: str, subfolder: Optional[str] = None, torch_dtype: Optional[torch.dtype] = None,
                          strict: bool = False, **overrides):
    """cls(**config.json) + load_state_dict.  Unknown config keys are dropped, like diffusers' ConfigMixin does; keys the
    checkpoint has but the mirror does not (e.g. the sincos `pos_embedding` buffer of non-RoPE models) are ignored unless
    `strict`."""
    cfg = load_config(root, subfolder)
    cfg.update(overrides)
    params = inspect.signature(cls.__init__).parameters
    if not any(p.kind == p.VAR_KEYWORD for p in params.values()):
        cfg = {k: v for k, v in cfg.items() if k in params}
    model = cls(**cfg)
    sd = load_state_dict(root, subfolder, torch_dtype)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if strict and (missing or unexpected):
        raise RuntimeError(f"{cls.__name__}: missing {missing[:5]} unexpected {unexpected[:5]}")
    if missing:
        real = [m for m in missing if "vip_" not in m]  # vip layers are loaded later by set_vip_layers
        if real:
            raise RuntimeError(f"{cls.__name__}: checkpoint lacks {len(real)} tensors, e.g. {real[:5]}")
    if torch_dtype is not None:
        model = model.to(torch_dtype)
    return model.eval()
