"""Import shim (golden-vector generation only) for the absent ``lightning`` package."""
from . import pytorch  # noqa: F401
