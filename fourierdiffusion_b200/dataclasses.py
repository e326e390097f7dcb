"""Module-path twin of the reference's `fdiff.utils.dataclasses` (the class itself lives in `batch.py`)."""
from .batch import DiffusableBatch  # noqa: F401
