class PeftAdapterMixin:
    pass


class CogVideoXLoraLoaderMixin:
    pass
