"""longvgen.schedulers (reference: longvgen/schedulers/__init__.py) -> tokensgen_b200.scheduler."""
from tokensgen_b200.scheduler import CogVideoXDPMScheduler  # noqa: F401
