"""Synthetic draw lists (frames of batches, tables and textures) for the tests and bench.py — input
generators, not part of the backend: `webrender_b200/` never imports this package."""
