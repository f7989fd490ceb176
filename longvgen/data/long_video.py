"""longvgen.data.long_video.load_video (reference :28-76) -> tokensgen_b200.video_io."""
from tokensgen_b200.video_io import load_video  # noqa: F401
