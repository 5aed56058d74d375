"""C4 (BASELINE.json configs[3]: RMSD pruning of a 200k-conformer ensemble sharded over the GPUs of one box), one
process per GPU over NCCL:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c4_sharded.py [n]
Rank 0 first prunes the ensemble alone (the single-GPU time and mask), then all ranks run the device-resident sharded
driver (fc_prune_sharded_dev) and the host-staged one.  Prints one JSON object (rank 0); times are maxima over the ranks."""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from firecode_b200 import dist as fdist
    from firecode_b200 import pruner, synthetic

    def tmax(seconds):
        t = torch.tensor([seconds], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rng = np.random.default_rng(synthetic.SEED + 4)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, n, 120, n // 100)
    out = {"n_gpus": world, "n": n}
    single_mask = None
    if rank == 0:
        for rep in range(3):
            t0 = time.perf_counter()
            _, single_mask = pruner.prune_by_rmsd(structures, atoms, 0.5)
            dt = time.perf_counter() - t0
        out["single_gpu"] = {"seconds": dt, "kept": int(single_mask.sum()), "library_ms": pruner.last_report.wall_ms}
        for rep in range(3):
            t0 = time.perf_counter()
            pruner.prune_by_rmsd(structures, atoms, 0.5, want_structures=False)
            out["single_gpu"]["seconds_mask_only"] = time.perf_counter() - t0
    dist.barrier()
    for label, staged in (("sharded_device_gather", False), ("sharded_host_staged", True)):
        times = []
        for rep in range(4):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            kept, mask = fdist.prune_sharded(structures, atoms, "rmsd", force_shard=True, host_staged=staged, max_rmsd=0.5)
            times.append(tmax(time.perf_counter() - t0))
        lib_ms = tmax(pruner.last_report.wall_ms)
        same = bool(np.array_equal(mask, single_mask)) if rank == 0 else True
        ok = torch.tensor([1.0 if same else 0.0], device=dev, dtype=torch.float64)
        sums = torch.tensor([float(np.flatnonzero(mask).sum())], device=dev, dtype=torch.float64)
        lo, hi = sums.clone(), sums.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        times_m = []
        for rep in range(3):   # the mask alone: no rank writes its copy of structures[mask] (232 MB each, one shared host memory)
            dist.barrier()
            t0 = time.perf_counter()
            _, mask_m = fdist.prune_sharded(structures, atoms, "rmsd", force_shard=True, host_staged=staged, max_rmsd=0.5,
                                            want_structures=False)
            times_m.append(tmax(time.perf_counter() - t0))
        assert np.array_equal(mask_m, mask)
        out[label] = {"seconds": min(times[1:]), "seconds_each": times, "seconds_mask_only": min(times_m[1:]),
                      "library_ms_max": lib_ms, "kept": int(mask.sum()),
                      "mask_equals_single_gpu": same, "all_ranks_same_mask": bool(lo.item() == hi.item())}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    main()
