"""Synthetic ensembles and pose sets of the BASELINE.json sizes.

Pure numpy, seeded, independent of the reference tree: used by tests/, bench.py and smoke() to
build identical inputs for the CUDA path and the CPU oracle.
"""

from __future__ import annotations

import numpy as np

SEED = 20261018


def random_rotations(rng, n):
    """(n,3,3) uniformly random rotation matrices from unit quaternions (x, y, z, w)."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    rot = np.empty((n, 3, 3))
    rot[:, 0, 0] = 1 - 2 * (y * y + z * z)
    rot[:, 0, 1] = 2 * (x * y - z * w)
    rot[:, 0, 2] = 2 * (x * z + y * w)
    rot[:, 1, 0] = 2 * (x * y + z * w)
    rot[:, 1, 1] = 1 - 2 * (x * x + z * z)
    rot[:, 1, 2] = 2 * (y * z - x * w)
    rot[:, 2, 0] = 2 * (x * z - y * w)
    rot[:, 2, 1] = 2 * (y * z + x * w)
    rot[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return rot


def molecule_cloud(rng, n_atoms, heavy_fraction=0.6):
    """Molecule-like point cloud: self-avoiding branched random walk of heavy atoms (1.5 A bonds)
    decorated with hydrogens (1.0-1.1 A), centred on its centroid.

    Returns (atoms (n,) str, coords (n,3) f64, bonds list[(i,j)], parent (n,) int)."""
    n_heavy = max(2, int(round(n_atoms * heavy_fraction)))
    n_heavy = min(n_heavy, n_atoms)
    coords = np.zeros((n_atoms, 3))
    parent = -np.ones(n_atoms, dtype=int)
    atoms = []
    bonds = []
    valence = np.zeros(n_atoms, dtype=int)

    def place(i, anchors, bond, min_d):
        for _ in range(2000):
            p = int(anchors[rng.integers(len(anchors))])
            if valence[p] >= 4:
                continue
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            cand = coords[p] + bond * v
            if i == 0 or np.min(np.linalg.norm(coords[:i] - cand, axis=1)) >= min_d:
                coords[i] = cand
                parent[i] = p
                valence[p] += 1
                valence[i] += 1
                bonds.append((p, i))
                return
        raise RuntimeError("could not place atom")

    atoms.append("C")
    for i in range(1, n_heavy):
        # branch with probability 0.3, otherwise extend the most recent atoms
        anchors = np.arange(i) if rng.random() < 0.3 else np.arange(max(0, i - 2), i)
        place(i, anchors, 1.5, 1.3)
        atoms.append(str(rng.choice(["C", "N", "O"], p=[0.7, 0.15, 0.15])))
    for i in range(n_heavy, n_atoms):
        place(i, np.arange(n_heavy), float(rng.uniform(0.95, 1.1)), 0.9)
        atoms.append("H")
    coords -= coords.mean(axis=0)
    return np.array(atoms), coords, bonds, parent


def _subtree_masks(n_atoms, bonds):
    """For every bond (p, c) of the tree: mask of atoms on the c side."""
    children = [[] for _ in range(n_atoms)]
    for p, c in bonds:
        children[p].append(c)
    masks = {}
    for p, c in bonds:
        mask = np.zeros(n_atoms, dtype=bool)
        stack = [c]
        while stack:
            k = stack.pop()
            mask[k] = True
            stack.extend(children[k])
        masks[(p, c)] = mask
    return masks


def rotate_about_axis(points, origin, axis, angle_rad):
    axis = axis / np.linalg.norm(axis)
    v = points - origin
    cos, sin = np.cos(angle_rad), np.sin(angle_rad)
    return origin + v * cos + np.cross(axis, v) * sin + np.outer(v @ axis, axis) * (1 - cos)


def conformer_ensemble(rng, n_conf, n_atoms, n_torsions=8, jitter=0.0):
    """n_conf conformers of one synthetic molecule: same topology, n_torsions rotatable bonds
    redrawn per conformer. Returns (atoms, coords (n_conf, n_atoms, 3), bonds, torsion_bonds)."""
    atoms, base, bonds, parent = molecule_cloud(rng, n_atoms)
    masks = _subtree_masks(n_atoms, bonds)
    # rotatable bonds: heavy-heavy bonds with a non-trivial subtree on the child side
    cands = [b for b in bonds if atoms[b[0]] != "H" and atoms[b[1]] != "H" and masks[b].sum() >= 2]
    picks = [cands[i] for i in rng.permutation(len(cands))[:n_torsions]] if cands else []
    out = np.empty((n_conf, n_atoms, 3))
    for c in range(n_conf):
        xyz = base.copy()
        if c > 0:
            for p, ch in picks:
                ang = rng.uniform(-np.pi, np.pi)
                m = masks[(p, ch)]
                xyz[m] = rotate_about_axis(xyz[m], xyz[ch], xyz[ch] - xyz[p], ang)
        if jitter > 0 and c > 0:
            xyz = xyz + rng.normal(scale=jitter, size=xyz.shape)
        out[c] = xyz
    out -= out.reshape(-1, 3).mean(axis=0)  # Hypermolecule centring: centroid over all conformers
    return atoms, out, bonds, picks


def radius_of_gyration(coords):
    c = coords - coords.mean(axis=0)
    return float(np.sqrt((c * c).sum() / len(c)))


def sweep_poses(rng, frag_a, frag_b, n_poses, shell=(-2.0, 4.0), dtype=np.float64):
    """C3 pose set: uniformly random rotations, translations uniform in direction with radius in
    Rg(A)+Rg(B)+shell. Returns xf (n_poses, 12): R row-major then t.  Values are rounded to
    float32 precision so that FP32 and FP64 consumers see identical inputs."""
    rot = random_rotations(rng, n_poses)
    direction = rng.normal(size=(n_poses, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    base = radius_of_gyration(frag_a) + radius_of_gyration(frag_b)
    radius = rng.uniform(base + shell[0], base + shell[1], size=(n_poses, 1))
    xf = np.concatenate([rot.reshape(n_poses, 9), direction * radius], axis=1)
    return np.ascontiguousarray(xf.astype(np.float32).astype(dtype))


def random_rigid(rng, coords):
    """Random rotation + translation of one structure."""
    rot = random_rotations(rng, 1)[0]
    return coords @ rot.T + rng.normal(scale=3.0, size=3)


def pruning_ensemble(rng, n_struct, n_atoms, n_basins, jitter=(0.05, 0.4), n_torsions=8):
    """C4-style ensemble: n_basins torsion-redraw conformers of one molecule, each copied with
    Gaussian per-atom jitter (sigma drawn per copy from `jitter`) and a random rigid motion,
    shuffled.  Returns (atoms, structures (n_struct, n_atoms, 3), basin id per structure)."""
    atoms, basins, _, _ = conformer_ensemble(rng, n_basins, n_atoms, n_torsions=n_torsions)
    basin_of = rng.integers(0, n_basins, size=n_struct)
    out = np.empty((n_struct, n_atoms, 3))
    for i, b in enumerate(basin_of):
        sigma = rng.uniform(*jitter)
        out[i] = random_rigid(rng, basins[b] + rng.normal(scale=sigma, size=(n_atoms, 3)))
    return atoms, out, basin_of
