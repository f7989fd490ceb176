"""longvgen.models.embeddings (reference :380-568, :571-707, :774-837, :920-984) -> tokensgen_b200.{transformer,rope}."""
from tokensgen_b200.rope import get_1d_rotary_pos_embed, get_3d_rotary_pos_embed, get_3d_rotary_pos_embed_v2  # noqa: F401
from tokensgen_b200.transformer import CogVideoXPatchEmbed, TimestepEmbedding, Timesteps  # noqa: F401
