"""Prophesee event / label file I/O with the reference's ``src/io`` API."""
from . import dat_events_tools, npy_events_tools  # noqa: F401
from .psee_loader import PSEELoader  # noqa: F401
