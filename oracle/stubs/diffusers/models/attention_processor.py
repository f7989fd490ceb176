class Attention:  # name only
    pass
