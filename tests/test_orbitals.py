"""Batched orbital centres (firecode_b200.orbitals) against the UNMODIFIED reactive-atom classes of the reference
(reactive_atoms_classes.py; Hypermolecule.compute_orbitals, hypermolecule_class.py:166-183): tests/golden/orbitals.npz
holds, for every class / variant, a duck-typed molecule with four conformers and the centres the reference computed
conformer by conformer (``python -m oracle.make_golden orbitals``); plus the centres of the reference's own fixture
molecules stored with the pivot tables (tests/golden/setup_rows.npz)."""

import os

import networkx as nx
import numpy as np
import pytest

from firecode_b200 import orbitals

HERE = os.path.dirname(os.path.abspath(__file__))


def _graph(n, bonds):
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from(map(tuple, bonds))
    return g


def test_orbital_centers_equal_reference_classes():
    gold = np.load(os.path.join(HERE, "golden", "orbitals.npz"), allow_pickle=False)
    names = [str(n) for n in gold["names"]]
    assert len(names) >= 15
    kinds_seen = set()
    for name in names:
        atoms = gold[f"{name}_atoms"]
        kind = str(gold[f"{name}_kinds"][0])
        dim = float(gold[f"{name}_orb_dim"])
        centers, subtype = orbitals.orbital_centers(
            kind, gold[f"{name}_coords"], atoms, _graph(len(atoms), gold[f"{name}_bonds"]), int(gold[f"{name}_index"]),
            orb_dim=orbitals.BOND_LENGTH if np.isnan(dim) else dim, reactive_indices=[int(r) for r in gold[f"{name}_reactive"]],
            sp3_sigmastar=bool(gold[f"{name}_sigmastar"]), sigmatropic=gold[f"{name}_sigmatropic"])
        ref = gold[f"{name}_centers"]
        assert centers.shape == ref.shape, name
        assert np.abs(centers - ref).max() < 1e-12, (name, np.abs(centers - ref).max())
        kinds_seen.add(kind.split(" (")[0])
        if subtype not in (None, "mixed"):
            assert f"({subtype})" in kind
    assert kinds_seen == set(orbitals.KINDS)


def test_single_bond_without_parameters_uses_the_bond_length(monkeypatch):
    """A single-bond centre of an element without an entry in orb_dim_dict sits one bond length beyond the atom."""
    gold = np.load(os.path.join(HERE, "golden", "orbitals.npz"), allow_pickle=False)
    name = "single_nodim"
    assert np.isnan(float(gold[f"{name}_orb_dim"]))
    monkeypatch.setattr(orbitals, "default_orb_dim", lambda symbol, kind: None)
    atoms = gold[f"{name}_atoms"]
    x = gold[f"{name}_coords"]
    centers, _ = orbitals.orbital_centers("Single Bond", x, atoms, _graph(len(atoms), gold[f"{name}_bonds"]),
                                          int(gold[f"{name}_index"]))
    i = int(gold[f"{name}_index"])
    assert np.allclose(np.linalg.norm(centers[:, 0] - x[:, i], axis=1), np.linalg.norm(x[:, i] - x[:, 0], axis=1))


def test_unsupported_kind_is_an_error_not_a_guess():
    g = _graph(3, [(0, 1), (0, 2)])
    with pytest.raises(NotImplementedError):
        orbitals.orbital_centers("sp", np.zeros((1, 3, 3)), ["C", "C", "C"], g, 0, orb_dim=1.0)


@pytest.mark.reference
def test_compute_orbitals_batch_equals_reference_on_its_fixtures():
    """The whole-molecule form against the UNMODIFIED Hypermolecule.compute_orbitals on the reference's own fixtures
    (every reactive atom of every molecule: sp3 / single-bond sigma-star pair, sp2, Ether, Ketone)."""
    from oracle import loader

    if not loader.reference_available():
        pytest.skip("reference tree not present")
    seen = set()
    for name in ("embed_string", "embed_cyclical", "embed_chelotropic", "embed_trimolecular"):
        with loader.embedder_from_dir(loader.fixture_dir(name), name + ".txt") as emb:
            for m, mol in enumerate(emb.objects):
                # orb_dim is the embedder's policy (parameter table, DIST pairings, per-embed adjustments such as
                # embedder.py:1079-1080, hypermolecule_class.py:239-240): it is read off the reference's centres wherever
                # the lobes sit at that distance from the atom; what is compared is the geometry of the lobes
                dims = orbitals.pairing_orb_dims(emb, m)
                for index, atom in mol.reactive_atoms_classes_dict[0].items():
                    length = float(np.linalg.norm(np.asarray(atom.center[0]) - mol.coords[0][int(index)]))
                    if repr(atom) == "Single Bond" and mol.sp3_sigmastar:
                        # (these lobes are not unit vectors times orb_dim: scale against the unit-orb_dim construction)
                        unit = orbitals.compute_orbitals_batch(mol, orb_dims={int(index): 1.0})[int(index)]
                        length /= float(np.linalg.norm(unit[0, 0] - mol.coords[0][int(index)]))
                    dims[int(index)] = length
                got = orbitals.compute_orbitals_batch(mol, orb_dims=dims)
                for index, atom in mol.reactive_atoms_classes_dict[0].items():
                    ref = np.array([np.asarray(mol.reactive_atoms_classes_dict[c][index].center, dtype=float)
                                    for c in range(len(mol.coords))])
                    assert got[int(index)].shape == ref.shape, (name, index, repr(atom))
                    assert np.abs(got[int(index)] - ref).max() < 1e-12, (name, index, repr(atom))
                    seen.add(repr(atom).split(" (")[0])
    assert {"sp3", "Single Bond", "sp2", "Ether", "Ketone"} <= seen
