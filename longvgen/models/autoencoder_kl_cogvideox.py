"""longvgen.models.autoencoder_kl_cogvideox (reference :38-1377) -> tokensgen_b200.vae."""
from tokensgen_b200.vae import (AutoencoderKLCogVideoX, CogVideoXCausalConv3d, CogVideoXDecoder3D,  # noqa: F401
                                CogVideoXDownBlock3D, CogVideoXEncoder3D, CogVideoXMidBlock3D, CogVideoXResnetBlock3D,
                                CogVideoXSafeConv3d, CogVideoXSpatialNorm3D, CogVideoXUpBlock3D)
