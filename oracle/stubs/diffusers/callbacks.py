"""Name-only: the reference pipelines import these for type hints (pipeline_cogvideox_mp_fifo.py:35)."""


class PipelineCallback:
    tensor_inputs = []


class MultiPipelineCallbacks:
    tensor_inputs = []
