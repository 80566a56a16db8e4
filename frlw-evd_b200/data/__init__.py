"""Drop-in twins of the reference's ``data/`` encoders (``data/sparse_ops.py``)."""
from . import sparse_ops  # noqa: F401
