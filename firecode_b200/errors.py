"""Exceptions that cross the drop-in boundary (reference: firecode/errors.py:23-51)."""


class ZeroCandidatesError(Exception):
    """Raised by the embed functions when no pose survives (embeds.py:147-154, 576-583, 741-748)."""


class TriangleError(Exception):
    """Raised by polygonize for impossible triangles (utils.py:275-276)."""


class FirecodeB200Error(RuntimeError):
    """The CUDA library is missing, failed to load, or a C-ABI call returned an error."""
