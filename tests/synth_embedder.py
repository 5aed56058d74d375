"""Re-export of the synthetic duck-typed embedder builder (lives in the package so bench.py and
__graft_entry__.smoke() can use it too)."""

from firecode_b200.synthetic_embedder import make_embedder  # noqa: F401
