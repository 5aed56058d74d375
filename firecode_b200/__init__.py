"""firecode_b200 -- B200-native (sm_100a) embedding screen for FIRECODE.

Drop-in for the hot path of ntampellini/FIRECODE: candidate-pose generation, compenetration
(clash) filtering, similarity pruning and torsion rotation, behind the reference's own function
names (``firecode/embeds.py``, ``firecode/utils.py``, ``firecode/algebra.py``,
``firecode/torsion_module.py`` and the ``prism_pruner`` pruning entry points it calls).

All arithmetic on the path runs in hand-written CUDA kernels reached through the C-ABI library
``csrc/libfirecode_b200.so`` (``include/firecode_b200.h``).  There is NO CPU fallback: importing the
package is cheap and GPU-free, but any compute call raises ``FirecodeB200Error`` when the library
or a CUDA device is missing.
"""

from .errors import FirecodeB200Error, TriangleError, ZeroCandidatesError  # noqa: F401

__version__ = "0.1.0"
