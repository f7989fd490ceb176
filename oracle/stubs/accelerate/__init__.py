"""Name-only stand-in: longvgen/fifo_sampling/__init__.py and the reference pipelines import accelerate unconditionally."""


class Accelerator:
    def __init__(self, *a, **k):
        raise NotImplementedError("accelerate is not on the reproduced path")
