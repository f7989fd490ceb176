"""Empty stand-in: the reference imports xformers unconditionally (attention_processor.py:19-20) but never calls it on this path."""
