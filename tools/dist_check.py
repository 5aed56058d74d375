"""Multi-GPU functional check over NCCL (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dist_check.py
Every sharded driver must return, on every rank, exactly what the single-GPU call returns."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from firecode_b200 import dist as fdist
    from firecode_b200 import embeds, problem, pruner, synthetic
    from firecode_b200.synthetic_embedder import make_embedder

    # clash sweep: bitmask all-gather
    rng = np.random.default_rng(1)
    _, a, _, _ = synthetic.molecule_cloud(rng, 60)
    _, b, _, _ = synthetic.molecule_cloud(rng, 50)
    xf = synthetic.sweep_poses(rng, a, b, 50001)
    from firecode_b200.clash import compenetration_check_batch

    single = compenetration_check_batch(a, b, xf, thresh=1.5).mask
    assert np.array_equal(fdist.clash_screen_sharded(a, b, xf, 1.5), single)
    # string embed: fingerprints all-gather + ordered keep-first sweep
    emb = make_embedder("string", 4, 24, seed=3, n_orb=2)
    ref = embeds.string_embed(make_embedder("string", 4, 24, seed=3, n_orb=2))
    assert np.array_equal(fdist.string_embed_sharded(emb), ref)
    # trimolecular embed: conformer-triple ranges
    kw = dict(n_mols=3, n_conf=[3, 2, 2], n_atoms=20, seed=9, n_reactive=2, n_orb=1)
    e1, e2 = make_embedder("cyclical", **kw), make_embedder("cyclical", **kw)
    ref = embeds.cyclical_embed(e1)
    got = fdist.cyclical_embed_sharded(e2)
    assert np.array_equal(got, ref) and np.array_equal(e2.constrained_indices, e1.constrained_indices)
    # bimolecular embed: group-table slices
    kw = dict(n_conf=4, n_atoms=30, seed=5, n_reactive=2, n_orb=2)
    e1, e2 = make_embedder("cyclical", **kw), make_embedder("cyclical", **kw)
    ref = embeds.cyclical_embed(e1)
    got = fdist.cyclical_embed_sharded(e2)
    assert np.array_equal(got, ref) and np.array_equal(e2.constrained_indices, e1.constrained_indices)
    # pruning: tiles dealt to the ranks, similar-pair lists all-gathered
    atoms, structures, _ = synthetic.pruning_ensemble(np.random.default_rng(7), 6000, 40, 300, jitter=(0.02, 0.4))
    _, m1 = pruner.prune_by_rmsd(structures, atoms, 0.5)
    _, mw = fdist.prune_sharded(structures, atoms, "rmsd", force_shard=True, max_rmsd=0.5)          # exchange on the devices
    _, mh = fdist.prune_sharded(structures, atoms, "rmsd", force_shard=True, host_staged=True, max_rmsd=0.5)
    none, mm = fdist.prune_sharded(structures, atoms, "rmsd", force_shard=True, max_rmsd=0.5, want_structures=False)
    assert np.array_equal(m1, mw) and np.array_equal(m1, mh) and np.array_equal(m1, mm) and none is None
    _, i1 = pruner.prune_by_moment_of_inertia(structures, atoms)
    _, iw = fdist.prune_sharded(structures, atoms, "moi", force_shard=True)
    assert np.array_equal(i1, iw)
    dist.barrier()
    if rank == 0:
        print(f"dist_check OK on {world} GPUs (clash, string, trimolecular, bimolecular, pruning)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
