"""ctypes binding of the C-ABI library (include/firecode_b200.h). Fails loudly; no fallback."""

from __future__ import annotations

import ctypes as C
import os
import threading

from .errors import FirecodeB200Error

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libfirecode_b200.so")

_lock = threading.Lock()
_lib = None

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
VP = C.c_void_p



class Tie(C.Structure):
    """struct fc_tie"""

    _fields_ = [("a", C.c_int64), ("b", C.c_int64), ("value", C.c_double), ("kind", C.c_int32),
                ("decision", C.c_int32)]


class StringProblemC(C.Structure):
    """struct fc_string_problem"""

    _fields_ = [
        ("coords1", VP), ("n_conf1", C.c_int32), ("n_atoms1", C.c_int32),
        ("coords2", VP), ("n_conf2", C.c_int32), ("n_atoms2", C.c_int32),
        ("centers1", VP), ("vecs1", VP), ("k1", C.c_int32),
        ("centers2", VP), ("vecs2", VP), ("k2", C.c_int32),
        ("angles", VP), ("n_angles", C.c_int32),
        ("quadruplets", VP), ("n_quads", C.c_int32),
        ("thresh", C.c_double), ("max_clashes", C.c_int32), ("rot_handedness", C.c_int32),
        ("tfd_thresh", C.c_double),
    ]


class CyclicalProblemC(C.Structure):
    """struct fc_cyclical_problem"""

    _fields_ = [
        ("n_mols", C.c_int32),
        ("coords", VP * 3), ("n_conf", C.c_int32 * 3), ("n_atoms", C.c_int32 * 3),
        ("reactive", VP * 3), ("n_reactive", C.c_int32 * 3),
        ("n_groups", C.c_int64),
        ("group_conf", VP), ("group_pivot", VP), ("group_mean", VP), ("group_vecs", VP),
        ("group_dirs", VP), ("group_ids", VP), ("n_pairs", C.c_int32),
        ("angles", VP), ("n_angles", C.c_int32),
        ("thresh", C.c_double), ("max_clashes", C.c_int32), ("rot_handedness", C.c_int32),
        ("rmsd_thresh", C.c_double),
    ]


class Cyclical3ProblemC(C.Structure):
    """struct fc_cyclical3_problem"""

    _fields_ = [
        ("coords", VP * 3), ("n_conf", C.c_int32 * 3), ("n_atoms", C.c_int32 * 3),
        ("reactive", VP * 3), ("n_reactive", C.c_int32 * 3),
        ("pivot_offsets", VP * 3), ("pivot_vec", VP * 3), ("pivot_mean", VP * 3), ("pivot_ids", VP * 3),
        ("ratoms0", VP * 3), ("n_ratoms0", C.c_int32 * 3),
        ("angles", VP), ("n_angles", C.c_int32),
        ("pairings", VP), ("n_pairings", C.c_int32),
        ("internal", VP), ("n_internal", C.c_int32),
        ("thresh", C.c_double), ("rot_handedness", C.c_int32), ("rmsd_thresh", C.c_double),
        ("conf_tuple_lo", C.c_int64), ("conf_tuple_hi", C.c_int64),
        ("flags", C.c_int32),
    ]


# int (*fc_allgather_fn)(const void* send, int64_t send_bytes, const void** recv, int64_t* recv_bytes, void* ctx)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, VP, C.c_int64, C.POINTER(VP), C.POINTER(C.c_int64), VP)

# int (*fc_allgather_dev_fn)(const void* send_dev, void* recv_dev, int64_t bytes, void* stream, void* ctx)
ALLGATHER_DEV_FN = C.CFUNCTYPE(C.c_int, VP, VP, C.c_int64, VP, VP)

# name -> (restype, argtypes); must list every symbol include/firecode_b200.h declares
SIGNATURES = {
    "fc_result_free": (None, [VP]),
    "fc_result_counts": (C.c_int, [VP, c_i64p]),
    "fc_result_status": (C.c_int, [VP, VP]),
    "fc_result_survivors": (C.c_int, [VP, VP]),
    "fc_result_fingerprints": (C.c_int, [VP, VP]),
    "fc_result_kept_indices": (C.c_int, [VP, VP]),
    "fc_result_kept_coords": (C.c_int, [VP, VP]),
    "fc_result_constrained": (C.c_int, [VP, VP]),
    "fc_result_ties": (C.c_int64, [VP, VP, C.c_int64]),
    "fc_cyclical_screen": (C.c_int, [VP, C.POINTER(VP)]),
    "fc_cyclical3_screen": (C.c_int, [VP, C.POINTER(VP)]),
    "fc_cyclical_groups": (C.c_int, [VP, VP, VP, VP, VP, VP, VP, VP, VP, C.c_double, VP, C.c_int32, VP, C.c_int32, C.c_int64,
                                     c_i64p, VP, VP, VP, VP, VP, VP]),
    "fc_result_groups": (C.c_int, [VP, VP, VP]),
    "fc_prune": (C.c_int, [VP, C.c_int64, C.c_int32, C.c_int32, VP, C.c_int32, VP, C.c_double, C.c_double,
                           C.c_double, VP, C.c_double, C.c_int32, C.c_int32, C.c_int32, VP, VP, VP,
                           C.c_int64, c_i64p]),
    "fc_prune_sharded": (C.c_int, [VP, C.c_int64, C.c_int32, C.c_int32, VP, C.c_int32, VP, C.c_double, C.c_double,
                                   C.c_double, VP, C.c_double, C.c_int32, C.c_int32, C.c_int32, VP, VP, VP,
                                   C.c_int64, c_i64p, C.c_int32, C.c_int32, ALLGATHER_FN, VP]),
    "fc_prune_sharded_dev": (C.c_int, [VP, C.c_int64, C.c_int32, C.c_int32, VP, C.c_int32, VP, C.c_double, C.c_double,
                                       C.c_double, VP, C.c_double, C.c_int32, C.c_int32, C.c_int32, VP, VP, VP,
                                       C.c_int64, c_i64p, C.c_int32, C.c_int32, ALLGATHER_DEV_FN, VP]),
    "fc_bond_graph_batch": (C.c_int, [VP, C.c_int64, C.c_int32, VP, C.c_double, VP, VP]),
    "fc_bond_delta_batch": (C.c_int, [VP, C.c_int64, C.c_int32, VP, C.c_double, VP, C.c_int32, VP, VP, VP]),
    "fc_tfd_fingerprints": (C.c_int, [VP, C.c_int64, C.c_int32, VP, C.c_int32, VP]),
    "fc_tfd_first_match": (C.c_int, [VP, C.c_int64, C.c_int32, VP, VP, C.c_int64, C.c_double, VP, VP, C.c_int64, c_i64p]),
    "fc_csearch_apply": (C.c_int, [VP, C.c_int32, C.c_int32, VP, C.c_int32, VP, VP, C.c_int64, C.c_double, C.c_int32,
                                   C.c_int32, VP, VP, VP]),
    "fc_torsion_scan": (C.c_int, [VP, C.c_int32, C.c_int32, VP, C.c_int32, VP, VP, C.c_int32, C.c_double,
                                  C.c_int32, C.c_int32, C.c_int32, VP, VP, VP]),
    "fc_string_n_poses": (C.c_int64, [VP]),
    "fc_string_screen": (C.c_int, [VP, C.POINTER(VP)]),
    "fc_string_stage1": (C.c_int, [VP, C.c_int64, C.c_int64, C.POINTER(VP)]),
    "fc_tfd_keepfirst": (C.c_int, [VP, VP, C.c_int64, C.c_int32, C.c_double, VP, VP, C.c_int64, c_i64p]),
    "fc_string_materialize": (C.c_int, [VP, VP, C.c_int64, VP]),
    "fc_last_error": (C.c_char_p, []),
    "fc_version": (C.c_int, []),
    "fc_device_count": (C.c_int, []),
    "fc_clash_tile_poses": (C.c_int, [C.c_int]),
    "fc_clash_screen_dev": (C.c_int, [VP, C.c_int, C.c_int, VP, C.c_int, C.c_int, VP, C.c_int64, VP,
                                      C.c_int64, C.c_double, C.c_int, C.c_int, VP, VP, VP, VP, VP,
                                      C.c_int64, C.c_int64, VP]),
    "fc_clash_prepare_dev": (C.c_int, [VP, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(VP), VP]),
    "fc_clash_screen_prepared_dev": (C.c_int, [VP, VP, VP, C.c_int, C.c_int, VP, C.c_int64, VP, C.c_int64, C.c_int,
                                               C.c_int, VP, VP, VP, VP, VP, C.c_int64, C.c_int64, VP]),
    "fc_clash_prep_free": (None, [VP, VP]),
    "fc_clash_cell_meta": (C.c_int, [VP, C.c_int, C.c_int, C.c_double, c_fp]),
    "fc_clash_batch": (C.c_int, [VP, C.c_int, C.c_int, VP, C.c_int, C.c_int, VP, C.c_int64, VP,
                                 C.c_int64, C.c_double, C.c_int, C.c_int, VP, VP, VP, VP, VP,
                                 C.c_int64]),
    "fc_clash_screen_ex_dev": (C.c_int, [VP, VP, VP, C.c_int, C.c_int, VP, C.c_int, C.c_int64, VP, C.c_int64, C.c_int,
                                         C.c_int, VP, VP, VP, VP, VP, VP, C.c_int64, C.c_int64, VP, VP]),
    "fc_clash_batch_pose7": (C.c_int, [VP, C.c_int, C.c_int, VP, C.c_int, C.c_int, VP, C.c_int64, VP, C.c_int64,
                                       C.c_double, C.c_int, C.c_int, VP, VP, VP, VP, VP, C.c_int64]),
    "fc_host_alloc": (VP, [C.c_int64, c_i32p]),
    "fc_host_free": (None, [VP]),
    "fc_clash_geometry": (C.c_int, [C.c_int, c_i32p]),
    "fc_clash_timing": (C.c_int, [C.c_int, c_dp, c_i64p]),
    "fc_pack_mask_dev": (C.c_int, [VP, C.c_int64, VP, VP]),
    "fc_rmsd_and_max_batch": (C.c_int, [VP, VP, C.c_int64, C.c_int32, C.c_int32, VP, VP]),
    "fc_rmsd_rot_corr_pairs": (C.c_int, [VP, C.c_int64, C.c_int32, VP, C.c_int32, VP, C.c_int32, VP, VP, VP, VP, C.c_int64,
                                         C.c_int32, C.c_int32, VP, VP, VP, VP]),
    "fc_self_clash_batch": (C.c_int, [VP, C.c_int64, C.c_int32, VP, C.c_double, VP, VP]),
    "fc_structure_clash_batch": (C.c_int, [VP, C.c_int64, C.c_int32, VP, C.c_int32, C.c_double, VP, VP]),
    "fc_fitness_batch": (C.c_int, [VP, C.c_int64, C.c_int32, VP, VP, C.c_int32, VP]),
    "fc_prune_plan": (C.c_int, [VP, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, VP, C.c_int64,
                                VP, VP, C.c_int64, VP, VP]),
    "fc_prune_timing": (C.c_int, [VP]),
    "fc_prune_tiles": (C.c_int, [VP]),
    "fc_prune_plan_segments": (C.c_int, [VP, C.c_int64, C.c_int64, C.c_int64, C.c_int32, VP, C.c_int64, c_i64p]),
    "fc_kabsch_host": (C.c_int, [VP, VP, VP]),
    "fc_xyz_format": (C.c_int, [VP, C.c_int32, VP, C.c_int64, C.c_int32, VP, VP, C.c_int64, c_i64p]),
    "fc_take_rows": (C.c_int, [VP, C.c_int64, VP, C.c_int64, VP, C.c_int64]),
    "fc_probe_fp32_peak": (C.c_int, [c_dp, c_dp, VP]),
}


def load(require_device: bool = False):
    """Return the loaded library (ctypes.CDLL). Raises FirecodeB200Error if it cannot be used."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise FirecodeB200Error(
                    f"CUDA library not built: {LIB_PATH} is missing. Run "
                    "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C "
                    "firecode_b200/csrc`). firecode_b200 has no CPU fallback."
                )
            try:
                lib = C.CDLL(LIB_PATH)
            except OSError as exc:  # pragma: no cover
                raise FirecodeB200Error(f"cannot load {LIB_PATH}: {exc}") from exc
            for name, (res, args) in SIGNATURES.items():
                try:
                    fn = getattr(lib, name)
                except AttributeError as exc:
                    raise FirecodeB200Error(f"{LIB_PATH} does not export {name}") from exc
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    if require_device and _lib.fc_device_count() <= 0:
        raise FirecodeB200Error(
            "no CUDA device visible: firecode_b200 runs only on a B200 (sm_100a) and has no CPU "
            f"fallback ({_lib.fc_last_error().decode()})"
        )
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().fc_last_error().decode(errors="replace")
        raise FirecodeB200Error(f"{what} failed (code {rc}): {msg}")
