"""longvgen.models.cogvideox_transformer_3d (reference :54-332, :335-770) -> tokensgen_b200.transformer."""
from tokensgen_b200.transformer import CogVideoXBlock, CogVideoXTransformer3DModel  # noqa: F401
