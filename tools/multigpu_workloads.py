"""C2 / C4 / C5 over N GPUs (one process per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multigpu_workloads.py
Prints one JSON object (rank 0).  Times are wall-clock maxima over the ranks after a warm-up call."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from firecode_b200 import dist as fdist
    from firecode_b200 import embeds, problem, pruner, synthetic
    from firecode_b200.synthetic_embedder import make_embedder

    def tmax(seconds):
        t = torch.tensor([seconds], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"n_gpus": world}
    # ---- C2: trimolecular cyclical embed, conformer triples sharded over the ranks (weak in groups per rank? no:
    #      strong scaling -- the full 125 000 triples are split)
    emb = make_embedder("cyclical", 50, 60, seed=synthetic.SEED + 2, n_mols=3, n_reactive=2, n_orb=1)
    prob = problem.cyclical_problem(emb)
    n_units = 50 ** 3
    lo, hi = fdist.shard_bounds(n_units, world, rank)
    for rep in range(2):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        poses, cons, r = embeds.cyclical3_screen(prob, conf_tuple_range=(lo, hi), want_status=False, want_coords=False)
        counts = fdist.all_gather_varlen(np.array([r.n_poses, r.n_kept, r.n_clash_pass], dtype=np.int64))
        dt = tmax(time.perf_counter() - t0)
    tot = counts.reshape(world, 3).sum(axis=0)
    out["C2_trimolecular"] = {"poses": int(tot[0]), "kept": int(tot[1]), "clash_pass": int(tot[2]), "seconds": dt,
                              "poses_per_s": float(tot[0] / dt), "scaling": "strong"}
    # ---- C4: RMSD pruning of 200 k conformers x 120 atoms, pair tiles dealt to the ranks, pair lists all-gathered
    rng = np.random.default_rng(synthetic.SEED + 4)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 200000, 120, 2000)
    out["C4_rmsd_pruning_200k"] = {}
    for label, force in (("sharded", True), ("default", False)):
        for rep in range(2):
            dist.barrier()
            t0 = time.perf_counter()
            kept, mask = fdist.prune_sharded(structures, atoms, "rmsd", force_shard=force, max_rmsd=0.5)
            dt = tmax(time.perf_counter() - t0)
        rep_ = pruner.last_report
        pairs = fdist.all_gather_varlen(np.array([rep_.pairs_tiled], dtype=np.int64))
        n_pairs = int(pairs.sum()) if force else int(pairs[0])
        out["C4_rmsd_pruning_200k"][label] = {
            "kept": int(mask.sum()), "pairs_evaluated": n_pairs, "seconds": dt, "rmsd_pairs_per_s": float(n_pairs / dt),
            "passes": rep_.passes, "mask_checksum": int(np.flatnonzero(mask).sum()),
            "note": "work items dealt to the ranks, similar-pair lists all-gathered every pass" if force else
                    "dist.prune_sharded default (PRUNE_SHARD_MIN_PAIRS decides between sharding and pruning the whole "
                    "ensemble on every rank)"}
    dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    main()
