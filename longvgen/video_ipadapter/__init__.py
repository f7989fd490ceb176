"""longvgen.video_ipadapter (reference: resampler.py:66-245) -> tokensgen_b200.resampler."""
from tokensgen_b200.resampler import Resampler  # noqa: F401
