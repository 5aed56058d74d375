"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the
symbols include/firecode_b200.h declares; the ctypes table covers all of them; nothing in the
product imports the oracle; compute calls fail loudly without a CUDA device."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "firecode_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from firecode_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.fc_version() >= 100


def test_struct_layouts_match_header_sizes(tmp_path):
    """sizeof of every struct crossing the boundary, as gcc lays the header out, equals ctypes'."""
    import subprocess

    from firecode_b200 import _lib

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "firecode_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(fc_tie), sizeof(fc_string_problem), sizeof(fc_cyclical_problem), sizeof(fc_cyclical3_problem));'
                   'return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.Tie), ctypes.sizeof(_lib.StringProblemC), ctypes.sizeof(_lib.CyclicalProblemC),
                     ctypes.sizeof(_lib.Cyclical3ProblemC)]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "firecode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # comments may cite the oracle; code must never import / load / execute it
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert not re.search(r"(import_module|__import__|CDLL|dlopen)\([^)]*oracle", text), f


def test_compute_fails_loudly_without_gpu():
    from firecode_b200 import _lib, clash
    from firecode_b200.errors import FirecodeB200Error

    if _lib.load().fc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(FirecodeB200Error):
        clash.compenetration_check_batch(np.zeros((3, 3)), np.ones((3, 3)), np.zeros((1, 12)), thresh=1.0)


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call."""
    from firecode_b200 import _lib

    lib = _lib.load()
    rc = lib.fc_prune(None, -1, 3, 0, None, 0, None, 0.5, 1.0, 0.01, None, 0.0, 1, 0, 20, None, None, None, 0, None)
    assert rc == 1 and b"bad sizes" in lib.fc_last_error()
    g = (ctypes.c_int32 * 4)()
    assert lib.fc_clash_geometry(150, g) == 0 and g[0] * g[1] >= 150 and g[3] % 32 == 0
    assert lib.fc_clash_tile_poses(150) == g[2]
