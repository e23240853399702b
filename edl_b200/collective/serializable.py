"""Base class for user-defined checkpoint state (reference: python/edl/collective/serializable.py:15-17)."""
from ..utils.json_serializable import SerializableBase  # noqa: F401
