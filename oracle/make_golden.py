"""TEST INFRASTRUCTURE: generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, through oracle.loader) on its own test fixtures and on seeded synthetic embedders.

    python -m oracle.make_golden

Each file stores the plain-array problem (so the GPU box, which has no reference tree, can replay
it) and the reference's outputs.  Shim conventions in force are stored alongside (PARITY UNPINNED
for everything that goes through oracle/prism_pruner).
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _conventions():
    loader.install()
    from prism_pruner import conventions

    return json.dumps({k: v for k, v in vars(conventions).items() if k.isupper()})


def pack_string_problem(prob):
    return {
        "coords1": prob.coords[0], "coords2": prob.coords[1], "centers1": prob.centers[0],
        "centers2": prob.centers[1], "vecs1": prob.vecs[0], "vecs2": prob.vecs[1],
        "angles": prob.angles, "quadruplets": prob.quadruplets, "thresh": np.float64(prob.thresh),
        "constrained": prob.constrained,
    }


def pack_cyclical_problem(prob):
    out = {"n_mols": np.int64(prob.n_mols), "angles": prob.angles, "thresh": np.float64(prob.thresh),
           "pairings": np.array(prob.pairings, dtype=np.int64).reshape(-1, 2),
           "internal_constraints": np.array(prob.internal_constraints, dtype=np.int64).reshape(-1, 2),
           "internal_constraints_is_array": np.bool_(prob.internal_constraints_is_array),
           "max_norm_delta": np.float64(prob.max_norm_delta), "ids": np.array(prob.ids, dtype=np.int64)}
    for m in range(prob.n_mols):
        out[f"ratoms0_{m}"] = np.asarray(prob.ratoms0[m], dtype=np.int64).reshape(-1, 2)
        out[f"coords{m}"] = prob.coords[m]
        out[f"reactive{m}"] = prob.reactive[m]
        counts = np.array([len(p) for p in prob.pivot_vec[m]], dtype=np.int64)
        out[f"pivot_counts{m}"] = counts
        out[f"pivot_vec{m}"] = np.concatenate(prob.pivot_vec[m]).reshape(-1, 3)
        out[f"pivot_mean{m}"] = np.concatenate(prob.pivot_mean[m]).reshape(-1, 3)
        out[f"pivot_ids{m}"] = np.concatenate(prob.pivot_ids[m]).reshape(-1, 2)
    return out


SYNTH_TRIMOL = {
    "synth_trimol_a": dict(n_conf=2, n_atoms=12, seed=3, n_mols=3, n_reactive=2, n_orb=1),
    "synth_trimol_b": dict(n_conf=[1, 2, 1], n_atoms=[16, 12, 14], seed=21, n_mols=3, n_reactive=1, n_orb=2),
}


def main():
    loader.install()
    from firecode.errors import ZeroCandidatesError

    from firecode_b200 import problem

    os.makedirs(GOLDEN, exist_ok=True)
    conv = _conventions()
    for name in ("embed_string", "embed_cyclical", "embed_chelotropic", "embed_trimolecular"):
        with loader.embedder_from_dir(loader.fixture_dir(name), name + ".txt") as emb:
            if emb.embed == "string":
                data = pack_string_problem(problem.string_problem(emb))
            else:
                data = pack_cyclical_problem(problem.cyclical_problem(emb))
            try:
                structures, constrained = loader.run_reference_embed(emb)
                zero = False
            except ZeroCandidatesError:
                n_tot = int(sum(emb.ids))
                structures, constrained, zero = np.zeros((0, n_tot, 3)), np.zeros((0, 0, 2), dtype=int), True
            np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), embed=np.array(emb.embed),
                                candidates=np.int64(emb.candidates), ref_structures=structures,
                                ref_constrained=np.asarray(constrained), zero_candidates=np.bool_(zero),
                                conventions=np.array(conv), **data)
            print(f"{name}: embed={emb.embed} candidates={emb.candidates} kept={len(structures)}"
                  f"{' (ZeroCandidatesError)' if zero else ''}")
    # seeded synthetic trimolecular embedders through the UNMODIFIED reference cyclical_embed
    # (the reference's own trimolecular fixture yields zero candidates)
    import contextlib
    import io

    from firecode.embeds import cyclical_embed
    from synth_embedder import make_embedder

    for name, kw in SYNTH_TRIMOL.items():
        emb = make_embedder("cyclical", **kw)
        data = pack_cyclical_problem(problem.cyclical_problem(emb))
        with contextlib.redirect_stdout(io.StringIO()):
            structures = cyclical_embed(emb)
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), embed=np.array("cyclical"),
                            ref_structures=structures, ref_constrained=np.asarray(emb.constrained_indices),
                            zero_candidates=np.bool_(False), conventions=np.array(conv), **data)
        print(f"{name}: kept={len(structures)}")


def tfd_cases():
    """Seeded ensembles for the TFD-pruning fixture: (name, n, n_atoms, n_basins, seed)."""
    return [("tfd_prune_a", 400, 20, 30, 1), ("tfd_prune_b", 1500, 16, 200, 2)]


def make_tfd_case(n, n_atoms, n_basins, seed):
    from firecode_b200 import synthetic

    rng = np.random.default_rng(seed)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, n, n_atoms, n_basins, jitter=(0.0, 0.08))
    quads = np.array([rng.choice(n_atoms, 4, replace=False) for _ in range(min(8, n_atoms - 3))], dtype=np.int64)
    return structures, quads


def main_tfd():
    """prune_conformers_tfd of the UNMODIFIED reference (firecode/torsion_module.py:957-1043)."""
    loader.install()
    from firecode.torsion_module import prune_conformers_tfd

    for name, n, n_atoms, n_basins, seed in tfd_cases():
        structures, quads = make_tfd_case(n, n_atoms, n_basins, seed)
        _, mask = prune_conformers_tfd(structures, quads)
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), params=np.array([n, n_atoms, n_basins, seed]),
                            quadruplets=quads, ref_mask=np.asarray(mask, dtype=bool),
                            checksum=np.float64(structures.sum()))
        print(f"{name}: kept {int(np.sum(mask))}/{n}")


class DuckTorsion:
    """The attributes the csearch drivers read from firecode.torsion_module.Torsion (torsion_module.py:139-145)."""

    def __init__(self, torsion, n_fold):
        self.torsion, self.n_fold = tuple(int(i) for i in torsion), int(n_fold)

    def get_angles(self):
        return {2: (0, 180), 3: (0, 120, 240), 4: (0, 90, 180, 270), 6: (0, 60, 120, 180, 240, 300)}.get(self.n_fold)


def make_csearch_case(n_atoms, n_tors, seed):
    """(atoms, coords, graph, [(torsion, n_fold)]) of a seeded synthetic molecule."""
    import networkx as nx

    from firecode_b200 import synthetic

    rng = np.random.default_rng(seed)
    atoms, cc, bonds, picks = synthetic.conformer_ensemble(rng, 1, n_atoms, n_torsions=n_tors)
    g = nx.Graph()
    g.add_nodes_from(range(n_atoms))
    g.add_edges_from(bonds)
    tors = []
    for p_, ch in picks:
        nb_p = [k for k in g.neighbors(p_) if k != ch]
        nb_c = [k for k in g.neighbors(ch) if k != p_]
        if nb_p and nb_c:
            tors.append(((nb_p[0], p_, ch, nb_c[0]), [3, 2, 4, 6][len(tors) % 4]))
    return atoms, cc[0], g, tors


CSEARCH_CASES = {"csearch_a": (30, 4, 5), "csearch_b": (45, 5, 8)}


def main_csearch():
    """random_csearch / clustered_csearch of the UNMODIFIED reference (torsion_module.py:436-571, 726-891)."""
    loader.install()
    from firecode.torsion_module import Torsion, clustered_csearch, random_csearch

    for name, (n_atoms, n_tors, seed) in CSEARCH_CASES.items():
        atoms, coords, g, tors = make_csearch_case(n_atoms, n_tors, seed)
        ref_t = []
        for t, nf in tors:
            obj = Torsion(*t)
            obj.n_fold = nf
            ref_t.append(obj)
        np.random.seed(seed)
        rnd = random_csearch(atoms, coords, ref_t, g, n_out=25, logfunction=None, interactive_print=False)
        np.random.seed(seed)
        clu = clustered_csearch(atoms, coords, ref_t, g, n=1000, n_out=100000, logfunction=None, interactive_print=False)
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), params=np.array([n_atoms, n_tors, seed]),
                            random=rnd, clustered=clu, checksum=np.float64(coords.sum()))
        print(f"{name}: random {rnd.shape}, clustered {clu.shape}")


REFINING_CASES = {"refining_a": (300, 3, 11), "refining_b": (600, 2, 12)}   # name -> (structures, fragments, seed)


def make_refining_case(n, n_frag, seed):
    """A duck-typed embedder state after the embed: n structures of n_frag rigid fragments.  Structure s belongs to
    basin s % n_basins (fragment offsets of the basin) with a small per-structure jitter, so that some structures
    have clashing fragments, the constrained distances scatter around their targets, and basins give similar pairs."""
    from firecode_b200 import synthetic

    rng = np.random.default_rng(seed)
    frags, ids = [], []
    for _ in range(n_frag):
        _, x, _, _ = synthetic.molecule_cloud(rng, int(rng.integers(8, 14)))
        frags.append(x)
        ids.append(len(x))
    atoms = np.array(["C"] * int(sum(ids)))
    offs = np.concatenate([[0], np.cumsum(ids)]).astype(int)
    n_basins = max(2, n // 6)
    basin_shift = rng.normal(size=(n_basins, n_frag, 3)) * 1.2
    structures = np.zeros((n, int(offs[-1]), 3))
    for s in range(n):
        for f in range(n_frag):
            structures[s, offs[f]:offs[f + 1]] = frags[f] + np.array([4.5 * f, 0.0, 0.0]) + basin_shift[s % n_basins, f] \
                + rng.normal(size=3) * 0.05
    pairs = np.array([[offs[f], offs[(f + 1) % n_frag] + 1] for f in range(n_frag)] + [[0, 2]])
    pairs = np.sort(pairs, axis=1)
    constrained = np.broadcast_to(pairs, (n,) + pairs.shape).copy()
    table = {chr(ord("a") + i): tuple(int(v) for v in pr) for i, pr in enumerate(pairs[:-1])}   # last pair: no letter
    dists = {lett: 4.0 + 0.5 * i for i, lett in enumerate(table)}
    return atoms, structures, ids, constrained, table, dists


def main_refining():
    """compenetration_refining / fitness_refining / similarity_refining of the UNMODIFIED reference
    (firecode/embedder.py:1954-2039, 1410-1514) called on a duck-typed embedder."""
    loader.install()
    from types import SimpleNamespace

    from firecode.embedder import RunEmbedding

    class Duck:
        apply_mask = RunEmbedding.apply_mask
        get_pairing_dists_from_constrained_indices = RunEmbedding.get_pairing_dists_from_constrained_indices
        zero_candidates_check = RunEmbedding.zero_candidates_check

        def log(self, *a, **k):
            pass

        def debuglog(self, *a, **k):
            pass

        def log_warnings(self):
            pass

    for name, (n, n_frag, seed) in REFINING_CASES.items():
        atoms, structures, ids, constrained, table, dists = make_refining_case(n, n_frag, seed)
        out = {}
        for step in ("compenetration", "fitness", "similarity"):
            d = Duck()
            d.embed, d.ids, d.atoms = "multiembed", ids, atoms
            d.options = SimpleNamespace(clash_thresh=1.5, max_clashes=2, rmsd=0.5)
            d.structures, d.constrained_indices = structures.copy(), constrained.copy()
            d.energies, d.exit_status = np.arange(n, dtype=float), np.zeros(n, dtype=bool)
            d.pairings_table, d.objects = table, [None] * n_frag
            d.get_pairing_dist_from_letter = lambda lett, dists=dists: dists[lett]
            tag = np.arange(n)
            d.tag = tag
            if step == "compenetration":
                d.constrained_indices = np.concatenate([constrained, np.broadcast_to(tag[:, None, None], (n, 1, 2))], axis=1)
                RunEmbedding.compenetration_refining(d)
                out[step] = d.constrained_indices[:, -1, 0].copy()
            elif step == "fitness":
                RunEmbedding.fitness_refining(d, threshold=3.0)
                out[step] = d.energies.astype(np.int64)
            else:
                RunEmbedding.similarity_refining(d, tfd=False, moi=True, rmsd=True)
                out[step] = d.energies.astype(np.int64)
            print(f"{name}: {step} kept {len(out[step])}/{n}")
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), params=np.array([n, n_frag, seed]),
                            kept_compenetration=out["compenetration"], kept_fitness=out["fitness"],
                            kept_similarity=out["similarity"], checksum=np.float64(structures.sum()))


def multiembed_arrangements(reactive1, reactive2):
    """The arrangement enumeration of multiembed_bifunctional (multiembed.py:39-48), same expressions."""
    loader.install()
    from itertools import permutations

    from firecode.utils import cartesian_product

    pairs = cartesian_product(reactive1, reactive2)
    return [((ix_1, ix_2), (iy_1, iy_2)) for ((ix_1, ix_2), (iy_1, iy_2)) in permutations(pairs, 2)
            if ix_1 != iy_1 and ix_2 != iy_2]


def main_multiembed():
    """firecode/tests/embed_multiembed through the UNMODIFIED reference: every arrangement of multiembed_bifunctional
    (multiembed.py:33-159) is run with the reference's own run_child_embedder, IN THIS PROCESS and in arrangement
    order (the reference farms them out to a process pool and collects them in completion order).  Stored per
    arrangement: the child's cyclical problem as plain arrays, the structures / constrained indices the child's
    generate_candidates produced, the target distances fitness_refining read, and what run_child_embedder returned
    after compenetration / fitness / similarity(rmsd=False) refining."""
    loader.install()
    import firecode.multiembed as mm
    from firecode.embedder import RunEmbedding

    from firecode_b200 import problem

    captured = {}
    orig = RunEmbedding.generate_candidates

    def spy(self):
        captured["problem"] = pack_cyclical_problem(problem.cyclical_problem(self))
        captured["embed"] = self.embed
        out = orig(self)
        captured["structures"] = np.array(self.structures, dtype=float)
        captured["constrained"] = np.asarray(self.constrained_indices).copy()
        captured["atoms"] = np.asarray(self.atoms)
        captured["table"] = {k: tuple(int(x) for x in v) for k, v in self.pairings_table.items()}
        captured["dists"] = {k: self.get_pairing_dist_from_letter(k) for k in self.pairings_table}
        captured["options"] = (float(self.options.clash_thresh), int(self.options.max_clashes), float(self.options.rmsd))
        return out

    RunEmbedding.generate_candidates = spy
    try:
        with loader.embedder_from_dir(loader.fixture_dir("embed_multiembed"), "embed_multiembed.txt") as emb:
            assert emb.embed == "multiembed"
            mol1, mol2 = emb.objects
            arrangements = multiembed_arrangements(mol1.reactive_indices, mol2.reactive_indices)
            out = {"n_arrangements": np.int64(len(arrangements)), "arrangements": np.array(arrangements, dtype=np.int64),
                   "conventions": np.array(_conventions())}
            total = 0
            for i, arr in enumerate(arrangements):
                captured.clear()
                with loader._quiet():
                    structures, constrained = mm.run_child_embedder(mol1.filename, mol2.filename,
                                                                    constrained_indices=np.array(arr), i=i, options=emb.options)
                structures = np.asarray(structures, dtype=float)
                if "problem" not in captured:
                    raise RuntimeError("child embed did not reach generate_candidates")
                for k, v in captured["problem"].items():
                    out[f"c{i}_{k}"] = v
                zero = "structures" not in captured
                n_tot = int(sum(captured["problem"]["ids"]))
                out[f"c{i}_embed"] = np.array(captured["embed"])
                out[f"c{i}_zero"] = np.bool_(zero)
                out[f"c{i}_ref_structures"] = captured.get("structures", np.zeros((0, n_tot, 3)))
                out[f"c{i}_ref_constrained"] = captured.get("constrained", np.zeros((0, 2, 2), dtype=np.int64))
                out[f"c{i}_final_structures"] = structures.reshape(-1, n_tot, 3) if structures.size else np.zeros((0, n_tot, 3))
                out[f"c{i}_final_constrained"] = np.asarray(constrained) if structures.size else np.zeros((0, 2, 2), dtype=np.int64)
                if not zero:
                    out[f"c{i}_table_keys"] = np.array(sorted(captured["table"]))
                    out[f"c{i}_table_pairs"] = np.array([captured["table"][k] for k in sorted(captured["table"])], dtype=np.int64)
                    out[f"c{i}_dists"] = np.array([np.nan if captured["dists"][k] is None else captured["dists"][k]
                                                   for k in sorted(captured["table"])], dtype=float)
                    out["atoms"] = captured["atoms"]
                    out["options"] = np.array(captured["options"])
                total += len(out[f"c{i}_final_structures"])
                print(f"multiembed child {i + 1}/{len(arrangements)} {arr}: embed={captured['embed']} "
                      f"generated {len(out[f'c{i}_ref_structures'])} -> {len(out[f'c{i}_final_structures'])} after refining")
            np.savez_compressed(os.path.join(GOLDEN, "embed_multiembed.npz"), **out)
            print(f"embed_multiembed: {len(arrangements)} arrangements, {total} structures in arrangement order")
    finally:
        RunEmbedding.generate_candidates = orig


def make_scramble_case(seed, n_frag=2, n_struct=24):
    """An assembly of n_frag small molecules (bond graphs from graphize of their rest geometry) and n_struct copies of
    it: untouched, slightly jittered (no bond changes), with one atom pulled away (a bond breaks) and with two fragments
    pushed together (bonds form); a few atoms are declared constrained (excluded)."""
    from prism_pruner.graph_manipulations import graphize

    from firecode_b200 import synthetic

    rng = np.random.default_rng(seed)
    frags, graphs, atoms = [], [], []
    for f in range(n_frag):
        sym, x, _, _ = synthetic.molecule_cloud(rng, int(rng.integers(7, 13)))
        frags.append(x + np.array([7.0 * f, 0.0, 0.0]))
        atoms += list(sym)
        graphs.append(graphize(sym, x))
    atoms = np.array(atoms)
    base = np.concatenate(frags)
    offs = np.concatenate([[0], np.cumsum([len(f) for f in frags])]).astype(int)
    structures = np.repeat(base[None], n_struct, axis=0)
    for s in range(n_struct):
        kind = s % 4
        if kind == 1:
            structures[s] += rng.normal(size=base.shape) * 0.02
        elif kind == 2:   # pull one atom of a random fragment away from the rest
            f = int(rng.integers(n_frag))
            a = int(rng.integers(offs[f], offs[f + 1]))
            structures[s, a] += rng.normal(size=3) * rng.uniform(0.3, 1.5)
        elif kind == 3:   # push fragment 1 towards fragment 0
            structures[s, offs[1]:offs[2]] -= np.array([rng.uniform(2.0, 5.5), 0.0, 0.0])
    excluded = sorted(int(v) for v in rng.choice(len(atoms), size=2, replace=False))
    return atoms, structures, graphs, excluded


def main_setup():
    """Setup-side rows of the UNMODIFIED reference: the pivot tables Embedder._set_pivots builds on
    the reference's own fixtures (embedder.py:902-987) and verdicts of utils.scramble_check / molecule_check
    (utils.py:341-400) on seeded assemblies."""
    loader.install()
    from firecode.utils import molecule_check, scramble_check

    out = {}
    n_mol = 0
    for name in ("embed_cyclical", "embed_chelotropic", "embed_trimolecular"):
        with loader.embedder_from_dir(loader.fixture_dir(name), name + ".txt") as emb:
            for m, mol in enumerate(emb.objects):
                key = f"piv{n_mol}"
                n_mol += 1
                n_conf = len(mol.coords)
                atoms_r = [list(mol.reactive_atoms_classes_dict[c].values()) for c in range(n_conf)]
                for a in range(len(atoms_r[0])):
                    out[f"{key}_centers{a}"] = np.array([np.asarray(atoms_r[c][a].center, dtype=float) for c in range(n_conf)])
                out[f"{key}_n_ratoms"] = np.int64(len(atoms_r[0]))
                out[f"{key}_suprafacial"] = np.bool_(bool(emb.options.suprafacial))
                out[f"{key}_sigmastar"] = np.bool_(bool(mol.sp3_sigmastar))
                out[f"{key}_name"] = np.array(f"{name}:{m}")
                for c in range(n_conf):
                    piv = mol.pivots[c]
                    out[f"{key}_c{c}_start"] = np.array([p.start for p in piv]).reshape(-1, 3)
                    out[f"{key}_c{c}_end"] = np.array([p.end for p in piv]).reshape(-1, 3)
                    out[f"{key}_c{c}_pivot"] = np.array([p.pivot for p in piv]).reshape(-1, 3)
                    out[f"{key}_c{c}_meanpoint"] = np.array([p.meanpoint for p in piv]).reshape(-1, 3)
                    out[f"{key}_c{c}_index"] = np.array([p.index_ for p in piv], dtype=np.int64).reshape(-1, 2)
                out[f"{key}_n_conf"] = np.int64(n_conf)
    out["n_pivot_mols"] = np.int64(n_mol)
    # scramble_check / molecule_check of the reference on seeded assemblies
    n_cases = 3
    for k in range(n_cases):
        atoms, structures, graphs, excluded = make_scramble_case(100 + k, n_frag=2 + (k == 2))
        out[f"scr{k}_atoms"] = np.array(atoms)
        out[f"scr{k}_structures"] = structures
        out[f"scr{k}_excluded"] = np.array(excluded, dtype=np.int64)
        out[f"scr{k}_n_frag"] = np.int64(len(graphs))
        for f, g in enumerate(graphs):
            out[f"scr{k}_graph{f}_edges"] = np.array(sorted(tuple(sorted(e)) for e in g.edges), dtype=np.int64).reshape(-1, 2)
            out[f"scr{k}_graph{f}_nodes"] = np.int64(len(g.nodes))
        for mx in (0, 1, 3):
            out[f"scr{k}_ok_max{mx}"] = np.array([scramble_check(atoms, s, excluded, graphs, max_newbonds=mx) for s in structures])
            out[f"scr{k}_mol_ok_max{mx}"] = np.array([molecule_check(atoms, structures[0], s, max_newbonds=mx) for s in structures])
    out["n_scramble_cases"] = np.int64(n_cases)
    np.savez_compressed(os.path.join(GOLDEN, "setup_rows.npz"), conventions=np.array(_conventions()), **out)
    print(f"setup_rows: {n_mol} pivot tables, {n_cases} scramble cases, verdicts "
          f"{[int(out[f'scr{k}_ok_max0'].sum()) for k in range(n_cases)]} of {len(structures)} pass at max_newbonds = 0")


class _DuckMol:
    """What a reactive-atom class reads from a Hypermolecule (reactive_atoms_classes.py)."""

    def __init__(self, atoms, edges, coords, reactive_indices, sp3_sigmastar, sigmatropic):
        import networkx as nx

        self.atoms = np.array(atoms)
        self.graph = nx.Graph()
        self.graph.add_nodes_from(range(len(atoms)))
        self.graph.add_edges_from(edges)
        self.coords = coords
        self.reactive_indices = list(reactive_indices)
        self.sp3_sigmastar = sp3_sigmastar
        self.sigmatropic = list(sigmatropic)


def orbital_cases():
    """(name, class name, atoms, bonds, rest geometry, reactive atom, reactive indices, sp3_sigmastar, sigmatropic per
    conformer) -- small molecules with idealised geometries; the conformers are jittered copies."""
    t = 1.0 / np.sqrt(3.0)
    methyl_x = [[0, 0, 0], [1.78, 0, 0], [-0.36, 1.03, 0], [-0.36, -0.51, 0.89], [-0.36, -0.51, -0.89]]
    ethene = [[0, 0, 0], [1.34, 0, 0], [-0.56, 0.93, 0], [-0.56, -0.93, 0], [1.90, 0.93, 0], [1.90, -0.93, 0]]
    ethene_bonds = [(0, 1), (0, 2), (0, 3), (1, 4), (1, 5)]
    cases = [
        ("single_plain", "Single", ["C", "Cl", "H", "H", "H"], [(0, 1), (0, 2), (0, 3), (0, 4)], methyl_x, 1, [1], False, None),
        ("single_nodim", "Single", ["C", "Si", "H", "H", "H"], [(0, 1), (0, 2), (0, 3), (0, 4)], methyl_x, 1, [1], False, None),
        ("single_sigmastar", "Single", ["C", "Cl", "H", "H", "H"], [(0, 1), (0, 2), (0, 3), (0, 4)], methyl_x, 1, [0, 1], True, None),
        ("sp3_leaving_group", "Sp3", ["C", "Cl", "H", "H", "H"], [(0, 1), (0, 2), (0, 3), (0, 4)], methyl_x, 0, [0], False, None),
        ("sp3_one_heavy", "Sp3", ["C", "C", "H", "H", "H"], [(0, 1), (0, 2), (0, 3), (0, 4)], methyl_x, 0, [0], False, None),
        ("sp3_ambiguous", "Sp3", ["C", "C", "C", "H", "H"], [(0, 1), (0, 2), (0, 3), (0, 4)], methyl_x, 0, [0], False, None),
        ("sp3_sigmastar", "Sp3", ["C", "Cl", "H", "H", "H"], [(0, 1), (0, 2), (0, 3), (0, 4)], methyl_x, 0, [0, 1], True, None),
        ("sp2", "Sp2", ["C", "C", "H", "H", "H", "H"], ethene_bonds, ethene, 0, [0, 1], False, None),
        ("ether", "Ether", ["O", "C", "C", "H"], [(0, 1), (0, 2), (1, 3)], [[0, 0, 0], [1.1, 0.8, 0], [-1.1, 0.8, 0], [1.9, 0.1, 0.3]],
         0, [0], False, None),
        ("ketone_sp2", "Ketone", ["O", "C", "C", "H"], [(0, 1), (1, 2), (1, 3)], [[0, 0, 0], [1.22, 0, 0], [1.95, 1.25, 0], [1.80, -0.93, 0.1]],
         0, [0], False, [False, False, False, False]),
        ("ketone_mixed", "Ketone", ["O", "C", "C", "H"], [(0, 1), (1, 2), (1, 3)], [[0, 0, 0], [1.22, 0, 0], [1.95, 1.25, 0], [1.80, -0.93, 0.1]],
         0, [0], False, [True, False, True, True]),
        ("ketone_ketene", "Ketone", ["O", "C", "C", "H", "H"], [(0, 1), (1, 2), (2, 3), (2, 4)],
         [[0, 0, 0], [1.16, 0, 0], [2.47, 0, 0], [3.03, 0.93, 0], [3.03, -0.93, 0]], 0, [0], False, [False] * 4),
        ("ketone_trilobe", "Ketone", ["O", "C", "H", "H", "C"], [(0, 1), (1, 2), (1, 3), (1, 4)],
         [[0, 0, 0], [1.40, 0, 0], [1.76, 1.03, 0], [1.76, -0.51, 0.89], [1.90, -0.72, -1.25]], 0, [0], False, [False] * 4),
        ("imine_lone_pair", "Imine", ["N", "C", "C", "H"], [(0, 1), (0, 2), (1, 3)], [[0, 0, 0], [1.27, 0.2, 0], [-0.8, 1.2, 0], [1.8, -0.7, 0]],
         0, [0], False, [False] * 4),
        ("imine_p", "Imine", ["N", "C", "C", "H"], [(0, 1), (0, 2), (1, 3)], [[0, 0, 0], [1.27, 0.2, 0], [-0.8, 1.2, 0], [1.8, -0.7, 0]],
         0, [0], False, [True] * 4),
    ]
    del t
    return cases


def main_orbitals():
    """Orbital centres by the UNMODIFIED reactive-atom classes (reactive_atoms_classes.py) on duck-typed molecules with
    four jittered conformers each; the batched builder firecode_b200.orbitals is pinned to these."""
    loader.install()
    import firecode.reactive_atoms_classes as rac
    from firecode.parameters import orb_dim_dict

    rng = np.random.default_rng(20261018)
    out = {}
    names = []
    for name, cls_name, atoms, bonds, rest, index, reactive, sigmastar, sigmatropic in orbital_cases():
        rest = np.asarray(rest, dtype=float)
        coords = np.array([rest + rng.normal(size=rest.shape) * 0.04 for _ in range(4)])
        sigmatropic = [False] * 4 if sigmatropic is None else sigmatropic
        mol = _DuckMol(atoms, bonds, coords, reactive, sigmastar, sigmatropic)
        centers, kinds = [], []
        for c in range(4):
            atom = getattr(rac, cls_name)()
            atom.init(mol, index, update=True, conf=c)
            centers.append(np.asarray(atom.center, dtype=float))
            kinds.append(repr(atom))
        kind = kinds[0].split(" (")[0]
        dim = orb_dim_dict.get(f"{atoms[index]} {kind}")
        if dim is None and kind != "Single Bond":
            dim = orb_dim_dict["Fallback"]
        out[f"{name}_atoms"] = np.array(atoms)
        out[f"{name}_bonds"] = np.array(bonds, dtype=np.int64)
        out[f"{name}_coords"] = coords
        out[f"{name}_index"] = np.int64(index)
        out[f"{name}_reactive"] = np.array(reactive, dtype=np.int64)
        out[f"{name}_sigmastar"] = np.bool_(sigmastar)
        out[f"{name}_sigmatropic"] = np.array(sigmatropic)
        out[f"{name}_kinds"] = np.array(kinds)
        out[f"{name}_orb_dim"] = np.float64(np.nan if dim is None else dim)
        out[f"{name}_centers"] = np.array(centers)
        names.append(name)
        print(f"orbitals {name}: {kinds[0]} -> centres {np.array(centers).shape}")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLDEN, "orbitals.npz"), conventions=np.array(_conventions()), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "orbitals":
        main_orbitals()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "multiembed":
        main_multiembed()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "setup":
        main_setup()
        sys.exit(0)
    main()
    main_tfd()
    main_csearch()
    main_refining()
    main_multiembed()
    main_setup()
    main_orbitals()
