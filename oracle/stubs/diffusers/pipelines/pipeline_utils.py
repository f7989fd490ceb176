"""The slice of diffusers.DiffusionPipeline the reference pipelines use (pipeline_cogvideox_mp_fifo.py:341-352,373,1183,1340):
module registration, the execution device, a progress bar and the offload hook no-op."""
import contextlib

import torch


class _Bar:
    def update(self, n=1):
        pass

    def set_description(self, *a, **k):
        pass

    def set_postfix(self, *a, **k):
        pass


class DiffusionPipeline:
    def __init__(self):
        self._modules_registered = []

    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
            self._modules_registered.append(k)

    @property
    def _execution_device(self):
        for name in getattr(self, "_modules_registered", []):
            m = getattr(self, name)
            if isinstance(m, torch.nn.Module):
                for p in m.parameters():
                    return p.device
        return torch.device("cpu")

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()

    def maybe_free_model_hooks(self):
        pass
