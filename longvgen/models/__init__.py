"""longvgen.models (reference: longvgen/models/__init__.py) -> tokensgen_b200 mirrors."""
from tokensgen_b200.transformer import CogVideoXTransformer3DModel  # noqa: F401
from tokensgen_b200.vae import AutoencoderKLCogVideoX  # noqa: F401
