"""prism_pruner.algebra restated (TEST INFRASTRUCTURE; see package docstring).

Call sites that fix the contracts: firecode/embeds.py:132,239-252,381-399,539,694;
firecode/utils.py:246; firecode/algebra.py:25; firecode/torsion_module.py:1075;
firecode/hypermolecule_class.py:66-74.
"""

from __future__ import annotations

import numpy as np

from . import conventions


def norm_of(vec):
    """Euclidean norm of a 3-vector."""
    return float(np.sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]))


def normalize(vec):
    """vec / |vec| (firecode/algebra.py:25, reactive_atoms_classes.py:80)."""
    vec = np.asarray(vec, dtype=float)
    return vec / np.linalg.norm(vec)


def vec_angle(v1, v2):
    """Angle between two vectors in DEGREES (compared with 90/180 at embeds.py:239)."""
    v1_u = normalize(v1)
    v2_u = normalize(v2)
    return float(np.arccos(np.clip(np.dot(v1_u, v2_u), -1.0, 1.0)) * 180.0 / np.pi)


def quat_to_mat(q):
    """Unit quaternion (x, y, z, w), scalar last -> 3x3 rotation matrix."""
    x, y, z, w = q
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    m = np.empty((3, 3))
    m[0, 0] = x2 - y2 - z2 + w2
    m[1, 0] = 2 * (xy + zw)
    m[2, 0] = 2 * (xz - yw)
    m[0, 1] = 2 * (xy - zw)
    m[1, 1] = -x2 + y2 - z2 + w2
    m[2, 1] = 2 * (yz + xw)
    m[0, 2] = 2 * (xz + yw)
    m[1, 2] = 2 * (yz - xw)
    m[2, 2] = -x2 - y2 + z2 + w2
    return m


def rot_mat_from_pointer(pointer, angle):
    """Rotation by ``angle`` degrees about ``pointer`` through a unit quaternion.

    Handedness is the unpinned switch conventions.ROT_HANDEDNESS.
    """
    pointer = np.asarray(pointer, dtype=float)
    assert pointer.shape[0] == 3
    half = conventions.ROT_HANDEDNESS * float(angle) * np.pi / 180.0 / 2.0
    axis = pointer / np.linalg.norm(pointer)
    s = np.sin(half)
    return quat_to_mat((axis[0] * s, axis[1] * s, axis[2] * s, np.cos(half)))


def dihedral(p):
    """Signed dihedral of four points in degrees, atan2 form, range (-180, 180]."""
    p0, p1, p2, p3 = (np.asarray(x, dtype=float) for x in p)
    b0 = -1.0 * (p1 - p0)
    b1 = p2 - p1
    b2 = p3 - p2
    b1 = b1 / np.linalg.norm(b1)
    v = b0 - np.dot(b0, b1) * b1
    w = b2 - np.dot(b2, b1) * b1
    x = np.dot(v, w)
    y = np.dot(np.cross(b1, v), w)
    return float(np.degrees(np.arctan2(y, x)))


def get_inertia_moments(coords, masses):
    """Three principal moments of inertia about the centre of mass, ascending."""
    coords = np.asarray(coords, dtype=float)
    masses = np.asarray(masses, dtype=float)
    com = (coords * masses[:, None]).sum(axis=0) / masses.sum()
    r = coords - com
    tensor = np.zeros((3, 3))
    for (x, y, z), m in zip(r, masses):
        tensor[0, 0] += m * (y * y + z * z)
        tensor[1, 1] += m * (x * x + z * z)
        tensor[2, 2] += m * (x * x + y * y)
        tensor[0, 1] -= m * x * y
        tensor[0, 2] -= m * x * z
        tensor[1, 2] -= m * y * z
    tensor[1, 0], tensor[2, 0], tensor[2, 1] = tensor[0, 1], tensor[0, 2], tensor[1, 2]
    return np.sort(np.linalg.eigvalsh(tensor))
