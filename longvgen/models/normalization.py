"""longvgen.models.normalization (reference :34-92, :426-488) -> tokensgen_b200.transformer."""
from tokensgen_b200.transformer import AdaLayerNorm, CogVideoXLayerNormZero, CogVideoXVIPLayerNormZero  # noqa: F401
