"""longvgen.pipeline (reference: longvgen/pipeline/__init__.py) -> tokensgen_b200 pipelines."""
from tokensgen_b200.pipeline import MPFIFOVideoIPAdapterCogVideoXPipeline  # noqa: F401
from tokensgen_b200.pipeline_t2to import LongVGenCogVideoXPipeline  # noqa: F401
