import logging as _logging
from dataclasses import dataclass

import torch

USE_PEFT_BACKEND = False


def is_torch_version(op, ver):
    from packaging import version
    import operator
    ops = {">": operator.gt, ">=": operator.ge, "<": operator.lt, "<=": operator.le, "==": operator.eq}
    return ops[op](version.parse(torch.__version__.split("+")[0]), version.parse(ver))


class _LoggingShim:
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


logging = _LoggingShim()


def scale_lora_layers(model, weight):
    pass


def unscale_lora_layers(model, weight=None):
    pass


def deprecate(*args, **kwargs):
    pass


class BaseOutput(dict):
    """dataclass-style output with attribute + tuple access."""

    def __post_init__(self):
        for k, v in self.__dict__.items():
            self[k] = v

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


def is_torch_xla_available():
    return False


def replace_example_docstring(doc):
    def deco(fn):
        return fn
    return deco
