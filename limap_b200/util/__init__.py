"""limap.util — only the I/O of the artefacts on either side of the hot path."""
from . import io  # noqa: F401
