from dataclasses import dataclass

import torch

from ...utils import BaseOutput
from ...utils.torch_utils import randn_tensor


@dataclass
class DecoderOutput(BaseOutput):
    sample: torch.Tensor = None


class DiagonalGaussianDistribution:
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None):
        s = randn_tensor(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * s

    def mode(self):
        return self.mean
