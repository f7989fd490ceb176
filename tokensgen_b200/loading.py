"""Checkpoint loading for the mirrors: the diffusers directory layout the reference loads with `from_pretrained`
(infer_cogvideo_mp_fifo.py:150-176): `<root>/<subfolder>/config.json` + `diffusion_pytorch_model*.safetensors` (single
file or sharded with an index json) or `.bin`/`.pt`.  The module tree is built on the meta device and the checkpoint tensors
are assigned to it, read straight to the target device in the requested dtype; nothing is pickled between processes
(SURVEY §8-f4)."""
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


def load_state_dict(root: str, subfolder: Optional[str] = None, dtype: Optional[torch.dtype] = None,
                    device: Optional[Any] = None) -> Dict[str, torch.Tensor]:
    """`device`: where the tensors land.  safetensors shards are read straight to it (no host staging copy of a 11 GB
    transformer); pickled blobs go through `map_location`."""
    dev = str(torch.device(device)) if device is not None else "cpu"
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
            sd.update(load_file(os.path.join(d, s), device=dev))
    else:
        blobs = [f for f in files if f.endswith((".bin", ".pt", ".pth")) and "optimizer" not in f]
        if not blobs:
            raise FileNotFoundError(f"no weights (*.safetensors / *.bin / *.pt) under {d}")
        for b in blobs:
            sd.update(torch.load(os.path.join(d, b), map_location=dev, weights_only=True))
    if dtype is not None:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    return sd


def build_from_pretrained(cls, root: str, subfolder: Optional[str] = None, torch_dtype: Optional[torch.dtype] = None,
                          strict: bool = False, device: Optional[Any] = None, **overrides):
    """cls(**config.json) + load_state_dict.  Unknown config keys are dropped, like diffusers' ConfigMixin does; keys the
    checkpoint has but the mirror does not (e.g. the sincos `pos_embedding` buffer of non-RoPE models) are ignored unless
    `strict`.

    The module tree is built on the `meta` device (no random initialisation of 5.6 G parameters, no fp32 host copy) and the
    checkpoint tensors are ASSIGNED to it, read straight to `device` (default: host) in `torch_dtype` — the reference
    constructs on the CPU, copies the checkpoint over the random weights and moves the result, per GPU, serially
    (infer_cogvideo_mp_fifo.py:150-183, 211-213)."""
    cfg = load_config(root, subfolder)
    cfg.update(overrides)
    params = inspect.signature(cls.__init__).parameters
    if not any(p.kind == p.VAR_KEYWORD for p in params.values()):
        cfg = {k: v for k, v in cfg.items() if k in params}
    with torch.device("meta"):
        model = cls(**cfg)
    sd = load_state_dict(root, subfolder, torch_dtype, device)
    own = model.state_dict()
    for k, v in list(sd.items()):       # assignment keeps the checkpoint tensor as is: check what copy_ would have checked
        if k in own and tuple(own[k].shape) != tuple(v.shape):
            raise RuntimeError(f"{cls.__name__}: size mismatch for {k}: checkpoint {tuple(v.shape)} vs model {tuple(own[k].shape)}")
    missing, unexpected = model.load_state_dict(sd, strict=False, assign=True)
    if strict and (missing or unexpected):
        raise RuntimeError(f"{cls.__name__}: missing {missing[:5]} unexpected {unexpected[:5]}")
    if missing:   # (vip layers are not part of the tree yet: set_vip_layers adds and loads them later)
        raise RuntimeError(f"{cls.__name__}: checkpoint lacks {len(missing)} tensors, e.g. {missing[:5]}")
    left = [n for n, t in list(model.named_parameters()) + list(model.named_buffers()) if t.is_meta]
    if left:
        raise RuntimeError(f"{cls.__name__}: {len(left)} tensors were not materialised by the checkpoint, e.g. {left[:5]}")
    return model.eval()
