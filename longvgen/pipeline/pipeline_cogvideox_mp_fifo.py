"""longvgen.pipeline.pipeline_cogvideox_mp_fifo (reference :268-1518) -> tokensgen_b200.pipeline."""
from tokensgen_b200.pipeline import (CogVideoXPipelineOutput, FIFOCogVideoXPipelineOutput,  # noqa: F401
                                     MPFIFOVideoIPAdapterCogVideoXPipeline, get_resize_crop_region_for_grid,
                                     retrieve_timesteps)
