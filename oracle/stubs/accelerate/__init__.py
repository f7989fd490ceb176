"""Name-only stand-in: longvgen/fifo_sampling/__init__.py imports the unused accelerate variant unconditionally."""
