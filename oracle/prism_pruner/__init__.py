"""TEST INFRASTRUCTURE ONLY -- numpy restatement ("shim") of the third-party package
``prism_pruner`` that FIRECODE imports for its low-level numerics.

prism_pruner (PyPI ``prism-pruner``, pinned 0.0.7 by /root/reference/pixi.lock:171,5054-5063)
is NOT vendored in /root/reference and cannot be installed offline, so its published algorithms
are restated here from the call sites in the reference (cited per function) and from recollection
of the upstream project.  PARITY UNPINNED: the reference's tests hold no golden vectors for any
symbol in this package (tests/test_suite.py:73-84 assert exit codes only).  Every convention that
could not be verified is a named module-level switch in ``oracle.prism_pruner.conventions``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (firecode_b200/) never does.
"""

__version__ = "0.0.7+shim"
