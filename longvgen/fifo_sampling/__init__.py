"""longvgen.fifo_sampling (reference: cogvideo_sampling_mp_fifo.py:27-395) -> tokensgen_b200.fifo."""
from tokensgen_b200.fifo import cogvideo_fifo_mp_v2  # noqa: F401
