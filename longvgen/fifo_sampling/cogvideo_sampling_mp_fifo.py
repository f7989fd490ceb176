"""longvgen.fifo_sampling.cogvideo_sampling_mp_fifo (reference :27-579) -> tokensgen_b200.fifo."""
from tokensgen_b200.fifo import cogvideo_fifo_mp_v2  # noqa: F401
from tokensgen_b200.pipeline import CogVideoXPipelineOutput  # noqa: F401
