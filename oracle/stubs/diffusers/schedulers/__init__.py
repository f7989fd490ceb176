class CogVideoXDDIMScheduler:  # name-only (type hint in the reference pipelines)
    pass
