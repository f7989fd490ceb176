"""yaml configs with attribute access — the subset of OmegaConf the reference CLI relies on (infer_cogvideo_mp_fifo.py:186-389:
`args.x`, `args.get("x", default)`, `dps.update(...)`, `inputs.pop("public")`, iteration over items).  omegaconf is not a
dependency of this repo; the yaml schema (config/infer/{edit,gen}.yaml of the reference) is unchanged."""
from __future__ import annotations

import copy
from typing import Any

import yaml


class Config(dict):
    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = wrap(v)

    def __deepcopy__(self, memo):
        return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def update(self, other=(), **kw):
        for k, v in dict(other or {}, **kw).items():
            self[k] = wrap(v)


def wrap(x: Any) -> Any:
    if isinstance(x, dict) and not isinstance(x, Config):
        return Config({k: wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x


def load(path: str) -> Config:
    with open(path) as f:
        return wrap(yaml.safe_load(f) or {})


def save(cfg: Config, path: str) -> None:
    def plain(x):
        if isinstance(x, dict):
            return {k: plain(v) for k, v in x.items()}
        if isinstance(x, list):
            return [plain(v) for v in x]
        return x
    with open(path, "w") as f:
        yaml.safe_dump(plain(cfg), f, sort_keys=False)
