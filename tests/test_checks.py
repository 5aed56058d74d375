"""Setup-side rows: pivot tables (Embedder._get_pivots / _set_pivots, embedder.py:902-987) and the
bond-graph post-filters scramble_check / molecule_check (utils.py:341-400).

CPU: the pivot builder and the oracle restatement of the two checks against tests/golden/setup_rows.npz, which
``python -m oracle.make_golden setup`` produced with the UNMODIFIED reference (its fixtures, its functions).
GPU: the batched CUDA checks (C-ABI fc_bond_graph_batch / fc_bond_delta_batch) against the oracle on the golden
assemblies and on seeded ones, empty batches and the single-structure forms included."""

import os

import networkx as nx
import numpy as np
import pytest

from firecode_b200 import checks
from oracle import port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "setup_rows.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN, allow_pickle=False)


def _graphs(gold, k):
    graphs = []
    for f in range(int(gold[f"scr{k}_n_frag"])):
        g = nx.Graph()
        g.add_nodes_from(range(int(gold[f"scr{k}_graph{f}_nodes"])))
        g.add_edges_from(map(tuple, gold[f"scr{k}_graph{f}_edges"]))
        graphs.append(g)
    return graphs


def test_pivot_tables_equal_reference(gold):
    n_mols = int(gold["n_pivot_mols"])
    assert n_mols >= 7
    seen_sigmastar = seen_single_atom = False
    for m in range(n_mols):
        key = f"piv{m}"
        centers = [gold[f"{key}_centers{a}"] for a in range(int(gold[f"{key}_n_ratoms"]))]
        sigmastar = bool(gold[f"{key}_sigmastar"])
        table, keep = checks.set_pivots(centers, suprafacial=bool(gold[f"{key}_suprafacial"]), sp3_sigmastar=sigmastar)
        seen_sigmastar |= sigmastar
        seen_single_atom |= len(centers) == 1
        for c in range(int(gold[f"{key}_n_conf"])):
            sel = keep[c]
            name = str(gold[f"{key}_name"])
            assert np.array_equal(table["index"][sel], gold[f"{key}_c{c}_index"]), name
            for what in ("start", "end", "pivot", "meanpoint"):
                assert np.array_equal(table[what][c][sel], gold[f"{key}_c{c}_{what}"]), (name, what)
    assert seen_sigmastar and seen_single_atom   # both filters / both branches are exercised by the reference's fixtures


def test_pivot_suprafacial_filter_keeps_the_two_shortest():
    rng = np.random.default_rng(5)
    c1 = rng.normal(size=(3, 2, 3))
    c2 = c1 + np.array([1.0, 0.0, 0.0]) + rng.normal(size=(3, 2, 3)) * 0.3
    table, keep = checks.set_pivots([c1, c2], suprafacial=True)
    for c in range(3):
        norms = np.linalg.norm(table["pivot"][c], axis=1)
        assert len(keep[c]) == 2 and set(keep[c]) == set(np.argsort(norms)[:2])
    empty = checks.get_pivots([c1, c2, c1])
    assert empty["pivot"].shape == (3, 0, 3)


def test_oracle_checks_equal_reference_verdicts(gold):
    for k in range(int(gold["n_scramble_cases"])):
        atoms, structures, excluded = gold[f"scr{k}_atoms"], gold[f"scr{k}_structures"], gold[f"scr{k}_excluded"]
        graphs = _graphs(gold, k)
        for mx in (0, 1, 3):
            got = np.array([port.scramble_check(atoms, s, excluded, graphs, mx) for s in structures])
            assert np.array_equal(got, gold[f"scr{k}_ok_max{mx}"])
            got = np.array([port.molecule_check(atoms, structures[0], s, mx) for s in structures])
            assert np.array_equal(got, gold[f"scr{k}_mol_ok_max{mx}"])
        assert 0 < gold[f"scr{k}_ok_max0"].sum() < len(structures)


def test_bits_round_trip():
    edges = [(0, 1), (1, 40), (5, 5), (33, 2)]
    bits = checks.bits_from_edges(41, edges)
    assert checks.edges_from_bits(bits) == [(0, 1), (1, 40), (2, 33)]
    g1, g2 = nx.path_graph(3), nx.path_graph(2)
    assert checks.edges_from_bits(checks.assembly_bits([g1, g2])) == [(0, 1), (1, 2), (3, 4)]


@pytest.mark.reference
def test_oracle_checks_equal_live_reference():
    from oracle import loader, make_golden

    if not loader.reference_available():
        pytest.skip("reference tree not present")
    loader.install()
    from firecode.utils import molecule_check, scramble_check

    atoms, structures, graphs, excluded = make_golden.make_scramble_case(7, n_frag=3, n_struct=16)
    for s in structures:
        for mx in (0, 2):
            assert port.scramble_check(atoms, s, excluded, graphs, mx) == scramble_check(atoms, s, excluded, graphs, max_newbonds=mx)
            assert port.molecule_check(atoms, structures[0], s, mx) == molecule_check(atoms, structures[0], s, max_newbonds=mx)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_scramble_and_molecule_check_match_golden(gpu, gold):
    for k in range(int(gold["n_scramble_cases"])):
        atoms, structures, excluded = gold[f"scr{k}_atoms"], gold[f"scr{k}_structures"], gold[f"scr{k}_excluded"]
        graphs = _graphs(gold, k)
        ok, delta, near = checks.scramble_check_batch(atoms, structures, excluded, graphs, 0, return_counts=True)
        assert near.sum() == 0
        ref_delta = np.array([len(port.scramble_delta(atoms, s, excluded, graphs)) for s in structures])
        assert np.array_equal(delta, ref_delta)
        for mx in (0, 1, 3):
            assert np.array_equal(checks.scramble_check_batch(atoms, structures, excluded, graphs, mx), gold[f"scr{k}_ok_max{mx}"])
            assert np.array_equal(checks.molecule_check_batch(atoms, structures[0], structures, mx), gold[f"scr{k}_mol_ok_max{mx}"])
        # single-structure forms with the reference's signatures, and the log line of a failing structure
        bad = int(np.flatnonzero(~gold[f"scr{k}_ok_max0"])[0])
        lines = []
        assert checks.scramble_check(atoms, structures[bad], excluded, graphs, 0, logfunction=lines.append, title="cand") is False
        found = port.scramble_delta(atoms, structures[bad], excluded, graphs)
        assert lines == [f"cand, scramble_check - found {len(found)} extra bonds: {found}"]
        assert checks.scramble_check(atoms, structures[0], excluded, graphs) is True
        assert checks.molecule_check(atoms, structures[0], structures[bad]) == bool(gold[f"scr{k}_mol_ok_max0"][bad])


@pytest.mark.gpu
@pytest.mark.parametrize("n_atoms,n_struct,seed", [(9, 40, 1), (33, 64, 2), (70, 500, 3), (200, 30, 4)])
def test_bond_graph_and_delta_match_oracle(gpu, n_atoms, n_struct, seed):
    """Random clouds around the bonding distance (many pairs near their limit in both directions), per-structure
    expected graphs (molecule_check against each structure's own earlier geometry), word boundaries at 32 / 64 atoms."""
    rng = np.random.default_rng(seed)
    atoms = rng.choice(np.array(["H", "C", "N", "O", "Cl"]), size=n_atoms)
    old = rng.normal(size=(n_struct, n_atoms, 3)) * (n_atoms ** (1 / 3)) * 0.9
    new = old + rng.normal(size=old.shape) * 0.15
    adj = checks.bond_graph_batch(atoms, new)
    for s in range(0, n_struct, max(1, n_struct // 12)):
        assert set(checks.edges_from_bits(adj[s])) == port.bond_set(atoms, new[s])
    ok, delta, near = checks.molecule_check_batch(atoms, old, new, 2, return_counts=True)
    for s in range(0, n_struct, max(1, n_struct // 12)):
        assert delta[s] == len(port.molecule_delta(atoms, old[s], new[s]))
        assert near[s] == port.bond_near_threshold(atoms, new[s])
    assert np.array_equal(ok, delta <= 2) and 0 < ok.sum() + 1


@pytest.mark.gpu
def test_checks_empty_batch(gpu):
    atoms = np.array(["C", "H", "H"])
    g = nx.path_graph(3)
    assert checks.scramble_check_batch(atoms, np.zeros((0, 3, 3)), [], [g]).shape == (0,)
    assert checks.bond_graph_batch(atoms, np.zeros((0, 3, 3))).shape == (0, 3, 1)
