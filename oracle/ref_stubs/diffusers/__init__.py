"""Import-only stand-in for `diffusers` (TEST INFRASTRUCTURE; see ../pytorch_lightning/__init__.py)."""
