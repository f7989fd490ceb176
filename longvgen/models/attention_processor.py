"""longvgen.models.attention_processor (reference :46-717, :1885-2155) -> tokensgen_b200.transformer."""
from tokensgen_b200.transformer import (Attention, CogVideoXAttnProcessor2_0,  # noqa: F401
                                        VideoIPAdapterCogVideoXAttnProcessor2_0)
