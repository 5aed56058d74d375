"""TEST INFRASTRUCTURE: import and run the UNMODIFIED reference from /root/reference.

Recipe verified in SURVEY.md 8(c): heavy optional dependencies (ase, rdkit, sella, matplotlib ...)
are replaced by MagicMock modules -- the embedding screen never touches them -- and the absent
``prism_pruner`` is provided by the shim in ``oracle/prism_pruner``.  /root/reference does not exist
on the GPU box: everything here is for the CPU container (golden-vector generation and pinning of
``oracle.port``); callers must check ``reference_available()`` first.
"""

from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.metadata
import io
import os
import shutil
import sys
import tempfile
from contextlib import contextmanager
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("FIRECODE_REFERENCE_ROOT", "/root/reference")
ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))

_MOCKED_TOPLEVEL = {
    "ase", "matplotlib", "sella", "rdkit", "prettytable", "InquirerPy", "mlfsm", "racerts",
    "openconf", "tblite", "xtb", "aimnet", "fairchem", "aimnet2calc", "torchani",
}


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _MOCKED_TOPLEVEL:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = MagicMock(name=spec.name)
        mod.__name__ = spec.name
        mod.__path__ = []
        mod.__spec__ = spec
        mod.__loader__ = self
        return mod

    def exec_module(self, module):
        return None


_installed = False


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "firecode"))


def install() -> None:
    """Make ``import firecode`` resolve to the unmodified reference (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.meta_path.insert(0, _MockFinder())
    # the shim must precede the reference on sys.path
    for p in (REFERENCE_ROOT, ORACLE_DIR):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    _orig_version = importlib.metadata.version

    def _version(name):
        if name == "firecode":
            return "2.0.4"
        return _orig_version(name)

    importlib.metadata.version = _version
    os.environ.setdefault("FIRECODE_FORCE_SINGLE_THREAD", "true")
    from firecode.__main__ import env_variables_handling

    cwd = os.getcwd()
    try:
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)  # env_variables_handling may read ./.firecoderc
            env_variables_handling()
    finally:
        os.chdir(cwd)
    _installed = True


class _TextSink(io.TextIOWrapper):
    def __init__(self):
        super().__init__(io.BytesIO(), encoding="utf-8", write_through=True)


@contextmanager
def _quiet():
    out, err = sys.stdout, sys.stderr
    sys.stdout, sys.stderr = _TextSink(), _TextSink()
    try:
        yield
    finally:
        sys.stdout, sys.stderr = out, err


def fixture_dir(name: str) -> str:
    return os.path.join(REFERENCE_ROOT, "firecode", "tests", name)


@contextmanager
def embedder_from_dir(src_dir: str, input_name: str, stamp: str = "oracle", quiet: bool = True):
    """Copy ``src_dir`` to a temp dir, build the reference's real Embedder on ``input_name`` and
    yield it with the cwd inside that temp dir (the Embedder chdirs and writes log files)."""
    install()
    from firecode.embedder import Embedder

    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="fc_oracle_")
    work = os.path.join(tmp, "w")
    shutil.copytree(src_dir, work)
    try:
        os.chdir(work)
        if quiet:
            with _quiet():
                emb = Embedder(input_name, stamp=stamp)
        else:
            emb = Embedder(input_name, stamp=stamp)
        yield emb
    finally:
        import logging

        logging.shutdown()
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference_embed(emb, quiet: bool = True):
    """Run the reference's own generate_candidates-equivalent embed function on a real Embedder.

    Returns (structures (P, N, 3) f64, constrained_indices). Raises ZeroCandidatesError as the
    reference does."""
    from firecode.embeds import cyclical_embed, string_embed

    fn = {"string": string_embed, "cyclical": cyclical_embed, "chelotropic": cyclical_embed}[emb.embed]
    if quiet:
        with _quiet():
            structures = fn(emb)
    else:
        structures = fn(emb)
    return structures, emb.constrained_indices
