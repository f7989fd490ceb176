import torch

from . import is_torch_version  # noqa: F401


def maybe_allow_in_graph(cls):
    return cls


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """torch.randn; a CPU generator draws on CPU and the result is moved (diffusers semantics)."""
    device = device or torch.device("cpu")
    gen_dev = generator.device if generator is not None and not isinstance(generator, (list, tuple)) else None
    if gen_dev is not None and gen_dev.type != torch.device(device).type and gen_dev.type == "cpu":
        return torch.randn(shape, generator=generator, device="cpu", dtype=dtype).to(device)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)
