from dataclasses import dataclass

import torch

from ..utils import BaseOutput


@dataclass
class Transformer2DModelOutput(BaseOutput):
    sample: torch.Tensor = None


@dataclass
class AutoencoderKLOutput(BaseOutput):
    latent_dist: object = None
