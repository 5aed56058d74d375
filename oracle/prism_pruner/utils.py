"""prism_pruner.utils restated (TEST INFRASTRUCTURE).

Call forms: rotate_dihedral at firecode/torsion_module.py:529,537,825,834; align_structures at
firecode/embedder.py:1703; flatten at torsion_module.py:475; time_to_string everywhere.
"""

from __future__ import annotations

import numpy as np

from . import conventions
from .algebra import rot_mat_from_pointer
from .graph_manipulations import d_min_bond
from .rmsd import get_alignment_matrix


def flatten(array, typefunc=float):
    """Flatten an arbitrarily nested iterable into a list of typefunc(x)."""
    out = []

    def rec(item):
        for el in item:
            if hasattr(el, "__iter__") and not isinstance(el, (str, bytes)):
                rec(el)
            else:
                out.append(typefunc(el))

    rec(array)
    return out


def time_to_string(total_time, verbose=False, digits=1):
    """Human-readable time span."""
    s = ""
    t = float(total_time)
    days, t = divmod(t, 86400.0)
    hours, t = divmod(t, 3600.0)
    minutes, t = divmod(t, 60.0)
    if days:
        s += f"{int(days)} days " if verbose else f"{int(days)}d "
    if hours:
        s += f"{int(hours)} hours " if verbose else f"{int(hours)}h "
    if minutes:
        s += f"{int(minutes)} minutes " if verbose else f"{int(minutes)}m "
    s += f"{round(t, digits)} seconds" if verbose else f"{round(t, digits)}s"
    return s


def rotate_dihedral(coords, dihedral, angle, mask=None, indices_to_be_moved=None):
    """Rotate the atoms selected by ``mask`` about the i2-i3 axis of ``dihedral`` by ``angle``
    degrees, pivoting on coords[i3]. Returns a new array."""
    coords = np.array(coords, dtype=float)
    _, i2, i3, _ = dihedral
    if mask is None:
        mask = np.zeros(len(coords), dtype=bool)
        mask[list(indices_to_be_moved)] = True
    axis = conventions.TORSION_AXIS_SIGN * (coords[i2] - coords[i3])
    mat = rot_mat_from_pointer(axis, angle)
    center = coords[i3]
    coords[mask] = (mat @ (coords[mask] - center).T).T + center
    return coords


def align_structures(structures, indices=None):
    """Centre every structure on the mean of ``indices`` and Kabsch-align it onto the first."""
    structures = np.array(structures, dtype=float)
    n_atoms = structures.shape[1]
    indices = slice(0, n_atoms) if indices is None else np.asarray(indices).ravel()
    out = np.empty_like(structures)
    ref_view = structures[0][indices]
    ref_center = ref_view.mean(axis=0)
    out[0] = structures[0] - ref_center
    ref = out[0][indices]
    for t in range(1, len(structures)):
        centred = structures[t] - structures[t][indices].mean(axis=0)
        try:
            rot = get_alignment_matrix(centred[indices], ref)
        except np.linalg.LinAlgError:
            rot = np.eye(3)
        out[t] = centred @ rot
    return out


def get_double_bonds_indices(coords, atoms):
    """Sorted index pairs of short C/N/O-C/N/O bonds (double-bond like)."""
    coords = np.asarray(coords, dtype=float)
    heavy = [i for i, a in enumerate(atoms) if str(a) in ("C", "N", "O")]
    out = []
    for a_i, i in enumerate(heavy):
        for j in heavy[a_i + 1:]:
            d = np.linalg.norm(coords[i] - coords[j])
            if d < d_min_bond(atoms[i], atoms[j], factor=0.94):
                out.append((i, j))
    return out
