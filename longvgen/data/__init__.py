"""longvgen.data: only the inference loader is on the reproduced path (training datasets are out of scope)."""
