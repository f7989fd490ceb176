"""longvgen.schedulers.scheduling_dpm_cogvideox (reference :136-541) -> tokensgen_b200.scheduler."""
from tokensgen_b200.scheduler import CogVideoXDPMScheduler  # noqa: F401
