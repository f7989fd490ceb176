"""longvgen.pipeline.pipeline_cogvideox_t2to (reference :297-912) -> tokensgen_b200.pipeline_t2to."""
from tokensgen_b200.pipeline import CogVideoXPipelineOutput  # noqa: F401
from tokensgen_b200.pipeline_t2to import LongVGenCogVideoXPipeline  # noqa: F401
