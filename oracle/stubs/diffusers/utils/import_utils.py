def is_torch_npu_available():
    return False


def is_xformers_available():
    return False
