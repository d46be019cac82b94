"""xrsdkit.tools-compatible helpers that sit on the intensity path."""
from . import peak_math  # noqa: F401
