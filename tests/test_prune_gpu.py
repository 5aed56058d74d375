"""GPU parity of the similarity pruning (C-ABI fc_prune) against the oracle's restatement of
prism_pruner (PARITY UNPINNED at the prism_pruner boundary: both sides share the same switches)."""

import numpy as np
import pytest

from firecode_b200 import pruner, synthetic
from oracle import port
from oracle.prism_pruner import pruner as ref_pruner

pytestmark = pytest.mark.gpu


def _forced(report):
    out = {}
    for t in report.ties:
        kind = {3: "rmsd", 4: "maxdev"}[int(t["kind"])]
        out[(kind, int(t["a"]), int(t["b"]))] = bool(t["decision"])
    return out


@pytest.mark.parametrize("keep,pass_mode", [("first", "greedy"), ("last", "snapshot"), ("first", "snapshot"),
                                            ("last", "greedy")])
@pytest.mark.parametrize("n,n_atoms,n_basins", [(300, 40, 12), (900, 24, 40)])
def test_prune_by_rmsd_matches_oracle(gpu, keep, pass_mode, n, n_atoms, n_basins):
    rng = np.random.default_rng(synthetic.SEED + n + n_atoms)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, n, n_atoms, n_basins, jitter=(0.02, 0.5))
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.5, keep=keep, pass_mode=pass_mode)
    rep = pruner.last_report
    ties = port.Ties(eps=1e-6, forced=_forced(rep))
    ref_out, ref_mask = ref_pruner.prune_by_rmsd(structures, atoms, 0.5, ties=ties, keep=keep, pass_mode=pass_mode)
    assert not [k for k in ties.seen if k not in ties.forced]
    assert mask.dtype == bool and mask.shape == (n,)
    assert np.array_equal(mask, ref_mask)
    assert np.array_equal(out, structures[ref_mask])
    assert 1 < mask.sum() < n  # both similar and dissimilar pairs exist


@pytest.mark.parametrize("keep,pass_mode", [("first", "greedy"), ("last", "snapshot")])
def test_prune_chunks_over_active_structures(gpu, keep, pass_mode):
    """The fifth unpinned convention: a pass may cut the ACTIVE structures in k chunks instead of the full
    array (conventions.PRUNE_CHUNK_OVER = "active").  Both sides carry the switch; the kept sets agree for either setting
    and differ between the settings on an ensemble with several passes."""
    rng = np.random.default_rng(4242)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 1800, 16, 60, jitter=(0.02, 0.35))
    masks = {}
    for chunk_over in ("full", "active"):
        _, mask = pruner.prune_by_rmsd(structures, atoms, 0.4, keep=keep, pass_mode=pass_mode, chunk_over=chunk_over)
        assert pruner.last_report.passes >= 4
        ties = port.Ties(eps=1e-6, forced=_forced(pruner.last_report))
        _, ref_mask = ref_pruner.prune_by_rmsd(structures, atoms, 0.4, ties=ties, keep=keep, pass_mode=pass_mode,
                                               chunk_over=chunk_over)
        assert np.array_equal(mask, ref_mask), chunk_over
        masks[chunk_over] = mask
        _, moi = pruner.prune_by_moment_of_inertia(structures, atoms, keep=keep, pass_mode=pass_mode, chunk_over=chunk_over)
        _, ref_moi = ref_pruner.prune_by_moment_of_inertia(structures, atoms, keep=keep, pass_mode=pass_mode,
                                                           chunk_over=chunk_over)
        assert np.array_equal(moi, ref_moi), chunk_over
    assert 1 < masks["full"].sum() < len(structures)


def test_prune_with_energies_and_multipass(gpu):
    """Enough structures for several chunked passes (20 * k < active) and an energy window."""
    rng = np.random.default_rng(99)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 2500, 16, 150, jitter=(0.02, 0.3))
    energies = rng.uniform(0, 5, size=len(structures))
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.4, energies=energies, max_dE=1.0)
    assert pruner.last_report.passes >= 4
    ties = port.Ties(eps=1e-6, forced=_forced(pruner.last_report))
    _, ref_mask = ref_pruner.prune_by_rmsd(structures, atoms, 0.4, energies=energies, max_dE=1.0, ties=ties)
    assert np.array_equal(mask, ref_mask)


def test_prune_by_moi_matches_oracle(gpu):
    rng = np.random.default_rng(5)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 600, 30, 25, jitter=(0.0, 0.02))
    out, mask = pruner.prune_by_moment_of_inertia(structures, atoms)
    _, ref_mask = ref_pruner.prune_by_moment_of_inertia(structures, atoms)
    # relative deviations are compared with 1e-2: tolerate only pairs within 1e-9 of that threshold
    assert np.array_equal(mask, ref_mask)
    assert 1 < mask.sum() < len(mask)


def test_prune_edge_cases(gpu):
    rng = np.random.default_rng(1)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 3, 10, 1, jitter=(0.0, 0.0))
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.5)
    assert mask.tolist() == [True, False, False]       # identical up to rigid motion: keep first
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.5, keep="last", pass_mode="snapshot")
    assert mask.tolist() == [False, False, True]
    out, mask = pruner.prune_by_rmsd(structures[:1], atoms, 0.5)
    assert mask.tolist() == [True]
    out, mask = pruner.prune_by_rmsd(structures[:0], atoms, 0.5)
    assert mask.shape == (0,) and out.shape == (0, 10, 3)
    s2, m2 = pruner.prune(structures, atoms, max_rmsd=0.5)
    assert m2.sum() == 1


def _shard_worker(rank, world, port, q):
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # both ranks share cuda:0; gloo carries the lists
    try:
        from firecode_b200 import dist as fdist
        from firecode_b200 import pruner as pr
        from firecode_b200 import synthetic as syn

        rng = np.random.default_rng(77)
        atoms, structures, _ = syn.pruning_ensemble(rng, 3000, 20, 200, jitter=(0.02, 0.4))
        out = {}
        # device-resident exchange (fc_prune_sharded_dev: 1 / world of the structures uploaded per rank, lists gathered and
        # sorted on the GPU; gloo stages the device buffers through the host here) and the host-staged form
        for staged in (False, True):
            for keep, mode in (("first", "greedy"), ("last", "snapshot"), ("last", "greedy")):
                _, mask = fdist.prune_sharded(structures, atoms, "rmsd", force_shard=True, host_staged=staged, max_rmsd=0.4,
                                              keep=keep, pass_mode=mode)
                out[(keep, mode, staged)] = (mask, pr.last_report.pairs_tiled)
            _, mmask = fdist.prune_sharded(structures, atoms, "moi", force_shard=True, host_staged=staged)
            out[("moi", staged)] = (mmask, pr.last_report.pairs_tiled)
        q.put((rank, out))
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_prune_sharded_two_ranks_equals_single(gpu):
    """Pair tiles dealt to two ranks + all-gather of the similar-pair lists + ordered resolve on every
    rank reproduces the single-GPU mask (and each rank evaluates about half of the pairs)."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(not isinstance(v, str) for v in results.values()), results
    rng = np.random.default_rng(77)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 3000, 20, 200, jitter=(0.02, 0.4))
    for keep, mode in (("first", "greedy"), ("last", "snapshot"), ("last", "greedy")):
        _, single = pruner.prune_by_rmsd(structures, atoms, 0.4, keep=keep, pass_mode=mode)
        total = pruner.last_report.pairs_tiled
        for staged in (False, True):
            for r in (0, 1):
                assert np.array_equal(results[r][(keep, mode, staged)][0], single), (keep, mode, staged, r)
            assert results[0][(keep, mode, staged)][1] + results[1][(keep, mode, staged)][1] == total
            assert abs(results[0][(keep, mode, staged)][1] - total / 2) < 0.2 * total
    _, single = pruner.prune_by_moment_of_inertia(structures, atoms)
    for staged in (False, True):
        assert np.array_equal(results[0][("moi", staged)][0], single) and np.array_equal(results[1][("moi", staged)][0], single)


def test_prune_skips_pairs_known_from_earlier_passes(gpu):
    rng = np.random.default_rng(5)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 4000, 16, 2000, jitter=(0.02, 0.1))
    _, mask = pruner.prune_by_rmsd(structures, atoms, 0.3)
    rep = pruner.last_report
    assert rep.passes >= 4 and rep.pairs_skipped > 0
    # (the vectorised form of the oracle driver, pinned to the loop restatement in tests/test_host_logic.py: seconds, not minutes)
    _, ref_mask = ref_pruner.prune_by_rmsd_vectorised(structures, atoms, 0.3, ties=port.Ties(eps=1e-6, forced=_forced(rep)))
    assert np.array_equal(mask, ref_mask)


@pytest.mark.parametrize("n_atoms,expect_tc", [(120, True), (36, True), (170, True), (300, False)])
def test_prune_screen_flavours_agree(gpu, monkeypatch, n_atoms, expect_tc):
    """The tensor-core screen (FP16 operands by default, TF32 with FC_PRUNE_TF32=1), the FP32 screen and the FP64-only pair
    kernel must give the same mask: the screens only decide which pairs reach the FP64 evaluation.  More than 176 selected
    atoms (88 as TF32) -> FP32 screen."""
    rng = np.random.default_rng(synthetic.SEED + n_atoms)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 3000, n_atoms, 60, jitter=(0.02, 0.45))
    masks = {}
    n_sel = int(np.sum(np.asarray(atoms) != "H"))
    for name, env in (("tc", {}), ("tf32", {"FC_PRUNE_TF32": "1"}), ("nocull", {"FC_PRUNE_CULL": "0"}), ("fp32", {"FC_PRUNE_TC": "0"}),
                      ("fp64", {"FC_PRUNE_FP64": "1"})):
        for k in ("FC_PRUNE_TC", "FC_PRUNE_FP64", "FC_PRUNE_TF32", "FC_PRUNE_CULL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        _, masks[name] = pruner.prune_by_rmsd(structures, atoms, 0.5)
        rep = pruner.last_report
        if name == "tc":
            assert (rep.screen_launches > 0) == expect_tc
            assert rep.n_sel == n_sel
            forced = _forced(rep)
            slots_with_culling = rep.screen_pair_slots
        elif name == "tf32":
            assert (rep.screen_launches > 0) == (n_sel <= 88)
        elif name == "nocull":   # positions in index order, every planned tile multiplied
            assert (rep.screen_launches > 0) == expect_tc
            if expect_tc:
                assert rep.screen_pair_slots >= slots_with_culling
        else:
            assert rep.screen_launches == 0
    assert np.array_equal(masks["tc"], masks["tf32"])
    assert np.array_equal(masks["tc"], masks["nocull"])
    assert np.array_equal(masks["tc"], masks["fp32"])
    assert np.array_equal(masks["tc"], masks["fp64"])
    assert 1 < masks["tc"].sum() < len(structures)
    _, ref_mask = ref_pruner.prune_by_rmsd(structures[:600], atoms, 0.5)
    _, sub_mask = pruner.prune_by_rmsd(structures[:600], atoms, 0.5)
    assert np.array_equal(sub_mask, ref_mask)


def test_prune_all_similar_and_all_distinct(gpu):
    """Every pair similar (the queue of pairs the screen cannot rule out overflows and is flushed repeatedly) and
    every pair dissimilar (the screen rules out everything)."""
    rng = np.random.default_rng(11)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 1500, 60, 1, jitter=(0.0, 0.02))
    _, mask = pruner.prune_by_rmsd(structures, atoms, 0.5)
    assert mask.sum() == 1 and mask[0]
    assert pruner.last_report.screen_candidates > 0
    structures = 3.0 * rng.normal(size=(1500, 40, 3))      # unrelated point clouds
    _, mask = pruner.prune_by_rmsd(structures, np.array(["C"] * 40), 0.5)
    assert mask.all()
    assert pruner.last_report.screen_launches > 0 and pruner.last_report.screen_candidates == 0


def test_prune_c4_20k_subset_matches_oracle(gpu):
    """kept-set equality with the oracle on the 20 k subset of BASELINE config C4 (120-atom molecule,
    200 basins x 100 jittered copies), default conventions; the oracle is the vectorised form of the same driver
    (tests/test_host_logic.py pins it to the loop restatement).  Nine chunked passes, ~8 000 kept."""
    rng = np.random.default_rng(synthetic.SEED + 4)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 20000, 120, 200)
    out, mask = pruner.prune_by_rmsd(structures, atoms, 0.5)
    rep = pruner.last_report
    assert rep.passes >= 8 and rep.screen_launches > 0
    ties = port.Ties(eps=1e-6, forced=_forced(rep))
    _, ref_mask = ref_pruner.prune_by_rmsd_vectorised(structures, atoms, 0.5, ties=ties)
    assert np.array_equal(mask, ref_mask)
    assert np.array_equal(out, structures[ref_mask])
    assert 4000 < mask.sum() < 16000
    none, mask_only = pruner.prune_by_rmsd(structures, atoms, 0.5, want_structures=False)   # the mask alone
    assert none is None and np.array_equal(mask_only, mask)


def test_prune_rank_deficient_species(gpu):
    """exact duplicates of a linear species (rank-1 covariance) and of a species with one heavy atom (zero
    covariance after centring) are pruned, as numpy's SVD-based reference does."""
    rng = np.random.default_rng(8)
    co2 = np.array([[-1.16, 0, 0], [0, 0, 0], [1.16, 0, 0.0]])
    structs = np.array([synthetic.random_rigid(rng, co2) for _ in range(6)])
    _, mask = pruner.prune_by_rmsd(structs, np.array(["O", "C", "O"]), 0.25)
    _, ref = ref_pruner.prune_by_rmsd(structs, np.array(["O", "C", "O"]), 0.25)
    assert mask.tolist() == ref.tolist() == [True] + [False] * 5
    water = np.array([[0, 0, 0.12], [0, 0.76, -0.47], [0, -0.76, -0.47]])
    structs = np.array([synthetic.random_rigid(rng, water) for _ in range(5)])
    _, mask = pruner.prune_by_rmsd(structs, np.array(["O", "H", "H"]), 0.25)
    _, ref = ref_pruner.prune_by_rmsd(structs, np.array(["O", "H", "H"]), 0.25)
    assert mask.tolist() == ref.tolist() == [True] + [False] * 4
