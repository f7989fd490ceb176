import logging


def get_logger(name, *a, **k):
    return logging.getLogger(name)
