class FromOriginalModelMixin:
    pass
