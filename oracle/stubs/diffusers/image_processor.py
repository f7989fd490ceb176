class IPAdapterMaskProcessor:
    pass
