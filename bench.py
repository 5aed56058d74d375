#!/usr/bin/env python
"""Benchmark of the firecode_b200 embedding screen (driver contract: one JSON line on stdout).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm
    python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm (host cores)

Workload (BASELINE.json configs[2], the one the metric "candidate poses screened/s ... at 1/2/4/8
B200" is quoted on): compenetration sweep of 10 M candidate poses of two 150-atom fragments PER GPU
(weak scaling; every rank screens its own seeded pose set, survivor bitmasks are all-gathered).
A step = one pass of the screen over that pose set:  table prep + (cell grid of fragment A) + FP32
screen kernel + FP64 recheck (+ bitmask pack + NCCL all-gather when N > 1).  The default screen is
the cell-list kernel; the all-pairs Gram-form kernel is timed on the same poses for
`roofline_allpairs` (FC_CLASH_MODE=0 forces it for the whole run).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ATOMS = 150
N_POSES = 10_000_000
THRESH = 1.5
FLOP_PER_PAIR = 8.0  # 3 sub + 3 mul + 2 add per atom pair (difference form)
EXEC_FLOP_PER_PAIR = 6.0  # what the Gram-form kernel executes: 3 FMA per atom pair
METRIC = "candidate poses screened/s"
UNIT = "poses/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


def make_fragments():
    from firecode_b200 import synthetic

    rng = np.random.default_rng(synthetic.SEED)
    _, a, _, _ = synthetic.molecule_cloud(rng, N_ATOMS)
    _, b, _, _ = synthetic.molecule_cloud(rng, N_ATOMS)
    return a, b


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle port of utils.py:544-551 looped per pose, as the
# reference does, on the host's physical cores (worker pool created OUTSIDE the timed window)
# ------------------------------------------------------------------------------------------------
_CPU_CTX = {}


def _cpu_init(a, b, thresh, use_reference):
    os.environ["OMP_NUM_THREADS"] = "1"
    _CPU_CTX["a"], _CPU_CTX["b"], _CPU_CTX["thresh"] = a, b, thresh
    fn = None
    if use_reference:
        try:
            from oracle import loader

            loader.install()
            from firecode.utils import compenetration_check as fn  # the UNMODIFIED reference function
        except Exception:
            fn = None
    if fn is None:
        from oracle import port

        fn = port.compenetration_check
    _CPU_CTX["fn"] = fn


def _cpu_worker(xf):
    from oracle import port

    a, b, thresh, fn = _CPU_CTX["a"], _CPU_CTX["b"], _CPU_CTX["thresh"], _CPU_CTX["fn"]
    ids = (len(a), len(b))
    passed = 0
    for p in range(len(xf)):
        pose = np.concatenate([a, port.place(b, xf[p])])  # get_embed, embeds.py:815-817
        passed += bool(fn(pose, ids=ids, thresh=thresh))
    return passed


def physical_cores():
    """Physical cores this process may use (hyperthread siblings counted once)."""
    allowed = sorted(os.sched_getaffinity(0))
    seen = set()
    for cpu in allowed:
        try:
            with open(f"/sys/devices/system/cpu/cpu{cpu}/topology/thread_siblings_list") as f:
                seen.add(f.read().strip())
        except OSError:
            seen.add(str(cpu))
    return max(1, len(seen)), len(allowed)


class CpuScreen:
    """Reference-style per-pose CPU screen on a pool of worker processes."""

    def __init__(self, cores, use_reference=False):
        import multiprocessing as mp

        self.cores = cores
        self.a, self.b = make_fragments()
        self.kind = "port"
        if use_reference and os.path.isdir("/root/reference/firecode"):
            self.kind = "reference"
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init,
                                                initargs=(self.a, self.b, THRESH, self.kind == "reference"))
        self.pool.map(_cpu_worker, [self.poses(8, seed=999)] * cores)  # workers imported and warm

    def poses(self, n, seed=0):
        from firecode_b200 import synthetic

        rng = np.random.default_rng(synthetic.SEED + 1000 + seed)
        return synthetic.sweep_poses(rng, self.a, self.b, n)

    def rate(self, n_sample, seed=0):
        xf = self.poses(n_sample, seed)
        parts = [xf[idx] for idx in np.array_split(np.arange(n_sample), self.cores) if len(idx)]
        t0 = time.perf_counter()
        out = self.pool.map(_cpu_worker, parts)
        wall = time.perf_counter() - t0
        return n_sample / wall, sum(out), wall

    def close(self):
        self.pool.close()
        self.pool.join()


WORKLOAD = "C3 compenetration sweep: 10M candidate poses of two 150-atom fragments per GPU, thresh 1.5 A"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores, logical = physical_cores()
    cpu = CpuScreen(cores)
    n_sample = 6000 * cores
    for i in range(args.warmup):
        cpu.rate(max(cores * 200, 200), seed=100 + i)
    rates, walls = [], []
    for i in range(args.steps):
        rate, _, wall = cpu.rate(n_sample, seed=i)
        rates.append(rate)
        walls.append(wall)
    cpu.close()
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls) * 1e3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "sample_poses_per_step": n_sample, "n_atoms": [N_ATOMS, N_ATOMS],
                   "note": "SAMPLED: each step screens a bounded sample of the same sweep (same fragments, same pose "
                           "distribution) with the reference's per-pose arithmetic on all physical host cores; the "
                           "worker pool is created outside the timed window"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "logical_cpus": logical, "kind": cpu.kind,
                         "sample": f"{n_sample} poses per step of the C3 sweep, per-pose "
                                   "get_embed + scipy cdist compenetration_check (oracle/port.py), "
                                   "one worker process per physical core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def _profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f), os.path.relpath(path, ROOT)
    return None, None


def run_cuda(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from firecode_b200 import _lib, clash, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lib = _lib.load(require_device=True)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_poses = args.poses
    a, b = make_fragments()
    a_dev = torch.from_numpy(a).to(dev)[None].contiguous()
    b_dev = torch.from_numpy(b).to(dev)[None].contiguous()

    # pose set of this rank, generated on the device with the same recipe as synthetic.sweep_poses:
    # a random quaternion and a translation on a shell, rounded to float32 (the compact pose, 28 B), and its
    # documented FP64 expansion (R | t, 96 B) -- the two forms the screen accepts
    def make_poses(n, seed):
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        q = torch.randn(n, 4, generator=gen, device=dev, dtype=torch.float64)
        q = q / q.norm(dim=1, keepdim=True)
        d = torch.randn(n, 3, generator=gen, device=dev, dtype=torch.float64)
        d = d / d.norm(dim=1, keepdim=True)
        base = synthetic.radius_of_gyration(a) + synthetic.radius_of_gyration(b)
        radius = base - 2.0 + 6.0 * torch.rand(n, 1, generator=gen, device=dev, dtype=torch.float64)
        p7 = torch.cat([q, d * radius], dim=1).float().contiguous()
        x, y, z, w = p7[:, :4].double().unbind(1)
        s = 2.0 / (((x * x + y * y) + z * z) + w * w)
        rot = torch.stack([1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
                           s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w),
                           s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)], dim=1)
        return p7, torch.cat([rot, p7[:, 4:].double()], dim=1).contiguous()

    pose7, xf = make_poses(n_poses, synthetic.SEED + 17 * rank)

    n_words = (n_poses + 31) // 32
    bits2 = [torch.empty(n_words, dtype=torch.int32, device=dev) for _ in range(2)]
    gathered2 = [torch.empty(world * n_words, dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else None
    near = (torch.zeros(4, dtype=torch.int32, device=dev), torch.zeros(4096, dtype=torch.int64, device=dev),
            torch.zeros(4096, dtype=torch.float64, device=dev))
    recheck_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    ev_bits = [torch.cuda.Event() for _ in range(2)]
    ev_gath = [torch.cuda.Event() for _ in range(2)]

    # kernels per step: A tables (Gram pairs, bounding box, cell grid), B tables (all-pairs layout, farthest-point
    # order), three levels of the cell-list screen (one kernel in all-pairs mode), FP64 recheck (+ pack in all-pairs mode)
    cell_mode = os.environ.get("FC_CLASH_MODE") != "0"
    launches_per_step = 9 if cell_mode else 6

    def screen(poses, fmt, n, bits_out, status_out=None):
        # one pass of the hot path: both fragments' tables, the screen, the FP64 recheck; survivors as a bitmask
        prep = clash.DevicePrep(a_dev, THRESH, want_cells=True)
        clash.screen_device_ex(prep, b_dev, poses[:n], fmt, bits_out=bits_out, status_out=status_out, near=near,
                               recheck_count=recheck_cnt)
        prep.free()

    step_i = [0]

    def step(poses=xf, fmt=clash.POSE_XF64, n=n_poses, words=n_words):
        k = step_i[0] & 1
        step_i[0] += 1
        if world > 1:
            torch.cuda.current_stream().wait_event(ev_gath[k])  # the all-gather that last read this buffer is done
        screen(poses, fmt, n, bits2[k])
        if world > 1:
            # the bitmask exchange of step k rides a side stream and overlaps the screen of step k + 1
            ev_bits[k].record()
            with torch.cuda.stream(side):
                side.wait_event(ev_bits[k])
                dist.all_gather_into_tensor(gathered2[k][: world * words], bits2[k][:words])
                ev_gath[k].record()

    def barrier():
        if world > 1:
            side.synchronize()
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, **kw):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(n_steps):
            step(**kw)
        if world > 1:
            torch.cuda.current_stream().wait_stream(side)  # the last exchange belongs to the timed region
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(1.0)  # nvidia-smi needs a moment before its first sample
    lib.fc_clash_timing(1, None, None)
    recheck_cnt.zero_()
    elapsed_ms = timed(args.steps)
    k_ms, k_n = C.c_double(0), C.c_int64(0)
    lib.fc_clash_timing(0, C.byref(k_ms), C.byref(k_n))
    clocks = sampler.stop() if rank == 0 else None
    n_recheck = int(recheck_cnt.item()) // max(1, args.steps)
    bits_cell = bits2[(step_i[0] - 1) & 1].clone()
    n_pass = int(sum(bin(int(w) & 0xffffffff).count("1") for w in bits_cell[:64].tolist()))  # spot value, full count below
    n_pass = int(torch.from_numpy(np.unpackbits(bits_cell.cpu().numpy().view(np.uint8))).sum().item())

    # ---- same screen on the compact poses (28 B per pose read instead of 96 B) ----------------------
    for _ in range(2):
        step(poses=pose7, fmt=clash.POSE_Q7)
    q7_ms = timed(max(1, min(args.steps, 10)), poses=pose7, fmt=clash.POSE_Q7) / max(1, min(args.steps, 10))
    bits_q7 = bits2[(step_i[0] - 1) & 1].clone()
    assert torch.equal(bits_q7, bits_cell), "compact-pose and f64-transform screens disagree"

    # ---- strong scaling (BASELINE.json configs[2] as written: 10 M poses IN TOTAL over the N GPUs) -----
    strong = None
    if world > 1:
        n_strong = (args.poses // world + 31) // 32 * 32
        w_strong = n_strong // 32
        for _ in range(3):
            step(n=n_strong, words=w_strong)
        s_steps = max(1, min(args.steps, 20))
        s_ms = timed(s_steps, n=n_strong, words=w_strong) / s_steps
        strong = {"scaling": "strong", "total_poses": n_strong * world, "poses_per_gpu": n_strong, "ms_per_step": s_ms,
                  "value": n_strong * world / (s_ms * 1e-3), "unit": UNIT, "steps": s_steps,
                  "note": "same step (tables + screen + recheck + bitmask all-gather on a side stream), 10 M poses split "
                          "over the ranks"}

    # ---- the all-pairs Gram-form kernel on the same poses (the formulation SURVEY.md 8d's FP32 roofline is
    #      written for); the default path above is the cell-list screen, which skips far atom pairs
    mode_default = os.environ.get("FC_CLASH_MODE")
    os.environ["FC_CLASH_MODE"] = "0"
    for _ in range(2):
        step()
    barrier()
    lib.fc_clash_timing(1, None, None)
    ap_steps = max(1, min(args.steps, 5))
    ap_ms_per_step = timed(ap_steps) / ap_steps
    ap_ms, ap_n = C.c_double(0), C.c_int64(0)
    lib.fc_clash_timing(0, C.byref(ap_ms), C.byref(ap_n))
    bits_ap = bits2[(step_i[0] - 1) & 1].clone()
    # mask EQUALITY of the two kernels on the full pose set (not only the pass count)
    assert torch.equal(bits_ap, bits_cell), "cell-list and all-pairs paths disagree"
    if mode_default is None:
        os.environ.pop("FC_CLASH_MODE", None)
    else:
        os.environ["FC_CLASH_MODE"] = mode_default
    cell_path = mode_default != "0"

    # ---- end to end through the public host API: pinned compact poses in, survivor bitmask out --------
    e2e_poses = min(n_poses, args.e2e_poses)
    p7_host = clash.pinned_empty((e2e_poses, 7), np.float32)
    p7_t = torch.from_numpy(p7_host)
    p7_t.copy_(pose7[:e2e_poses])
    bits_host = clash.pinned_empty(((e2e_poses + 31) // 32,), np.uint32)
    barrier()
    for _ in range(3):  # warm-up at the timed size: pinned staging buffer and pool reach their final size
        clash.compenetration_check_batch_pose7(a, b, p7_host, thresh=THRESH, bits_out=bits_host)
    barrier()
    e2e_steps = max(1, min(args.steps, 10))
    per_step = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        res = clash.compenetration_check_batch_pose7(a, b, p7_host, thresh=THRESH, bits_out=bits_host)
        per_step.append((time.perf_counter() - t1) * 1e3)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert np.array_equal(res.bits, bits_cell.cpu().numpy().view(np.uint32)[: len(res.bits)]) or e2e_poses != n_poses, \
        "e2e and device paths disagree"
    # raw pinned host -> device bandwidth of this box, for context: the e2e path moves 28 B per pose
    dst = torch.empty_like(pose7[:e2e_poses])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        dst.copy_(p7_t, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * p7_t.numel() * 4 / (time.perf_counter() - t0) / 1e9
    del dst
    e2e = {"value": world * e2e_poses / e2e_s, "unit": UNIT, "h2d_gbs_pinned_measured": h2d_gbs,
           "h2d_bound_poses_per_s": world * h2d_gbs * 1e9 / 28.0,
           "h2d_bytes_per_step": int(e2e_poses * 28 + (len(a) + len(b)) * 24),
           "d2h_bytes_per_step": int(res.bits.nbytes + 16), "poses_per_step": e2e_poses, "steps": e2e_steps,
           "ms_per_step_each": [round(x, 2) for x in per_step],
           "api": "firecode_b200.clash.compenetration_check_batch_pose7 -> C-ABI fc_clash_batch_pose7 (pinned host "
                  "buffers: 28 B per pose in, 1 bit per pose out)"}
    # the f64-transform form of the same API (96 B per pose in, status bytes out), for comparison with round 1
    xf_host = clash.pinned_empty((e2e_poses, 12), np.float64)
    torch.from_numpy(xf_host).copy_(xf[:e2e_poses])
    for _ in range(2):
        clash.compenetration_check_batch(a, b, xf_host, thresh=THRESH)
    t0 = time.perf_counter()
    for _ in range(3):
        res64 = clash.compenetration_check_batch(a, b, xf_host, thresh=THRESH)
    x64_s = (time.perf_counter() - t0) / 3
    assert res64.n_pass == res.n_pass or e2e_poses != n_poses
    e2e["xf64_api"] = {"value": world * e2e_poses / x64_s, "unit": UNIT, "h2d_bytes_per_step": int(e2e_poses * 96),
                       "d2h_bytes_per_step": int(e2e_poses),
                       "api": "firecode_b200.clash.compenetration_check_batch -> fc_clash_batch (f64 R|t, as round 1)"}
    del xf_host, p7_host

    # second half of the metric on several GPUs: C4 (configs[3], "... sharded over 8 GPUs"): every rank uploads 1 / world of
    # the structures, pair tiles dealt to the ranks, exchange steps on the devices over NCCL (fc_prune_sharded_dev)
    c4_sharded = None
    if world > 1 and not args.no_extras:
        from firecode_b200 import dist as fdist
        from firecode_b200 import pruner

        rng4 = np.random.default_rng(synthetic.SEED + 4)
        atoms4, structures4, _ = synthetic.pruning_ensemble(rng4, 200000, 120, 2000)
        times = []
        for rep_i in range(4):
            barrier()
            t0 = time.perf_counter()
            _, mask4 = fdist.prune_sharded(structures4, atoms4, "rmsd", force_shard=True, max_rmsd=0.5)
            t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        times_mask = []
        for rep_i in range(3):   # the kept mask alone: no rank writes its own 232 MB copy of structures[mask]
            barrier()
            t0 = time.perf_counter()
            _, mask4m = fdist.prune_sharded(structures4, atoms4, "rmsd", force_shard=True, max_rmsd=0.5, want_structures=False)
            t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times_mask.append(float(t.item()))
        assert np.array_equal(mask4m, mask4)
        lib_ms = torch.tensor([float(pruner.last_report.wall_ms)], device=dev, dtype=torch.float64)
        dist.all_reduce(lib_ms, op=dist.ReduceOp.MAX)
        pairs = torch.tensor([float(pruner.last_report.pairs_tiled)], device=dev, dtype=torch.float64)
        dist.all_reduce(pairs, op=dist.ReduceOp.SUM)
        chk = torch.tensor([float(np.flatnonzero(mask4).sum())], device=dev, dtype=torch.float64)
        chk_lo, chk_hi = chk.clone(), chk.clone()
        dist.all_reduce(chk_lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(chk_hi, op=dist.ReduceOp.MAX)
        sec = min(times[1:])
        c4_sharded = {"metric": "RMSD pairs/s", "value": float(pairs.item()) / sec, "unit": "pairs/s", "seconds": sec,
                      "seconds_each": [round(x, 4) for x in times], "n_gpus": world, "scaling": "strong",
                      "seconds_mask_only": min(times_mask[1:]), "pairs_per_s_mask_only": float(pairs.item()) / min(times_mask[1:]),
                      "library_ms_max_over_ranks": float(lib_ms.item()),
                      "note": "seconds = the drop-in call on every rank, each writing its own copy of structures[mask] (232 MB) "
                              "through the host memory the ranks share; seconds_mask_only = the same call with "
                              "want_structures=False (the mask is what the ranks need to agree on)",
                      "workload": "C4: prune_by_rmsd of 200 k conformers x 120 atoms through dist.prune_sharded (device all-gather)",
                      "pairs": int(pairs.item()), "kept": int(mask4.sum()), "mask_checksum": int(chk.item()),
                      "all_ranks_same_mask": bool(chk_lo.item() == chk_hi.item()),
                      "single_gpu_reference": "kept 80688, mask checksum in the N = 1 line (other_workloads.C4_rmsd_pruning_200k)"}
        del structures4, mask4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = elapsed_ms / args.steps
    value = world * n_poses / (ms_per_step * 1e-3)
    peaks, peak_kind = _peaks()
    kernel_ms = k_ms.value / max(1, k_n.value)
    pairs = float(n_poses) * N_ATOMS * N_ATOMS
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    clock_hz = peaks.get("sm_max_mhz", 1965.0) * 1e6
    peak_tf = sms * 128 * 2 * clock_hz / 1e12
    probe_tf, probe_ms = C.c_double(0), C.c_double(0)
    lib.fc_probe_fp32_peak(C.byref(probe_tf), C.byref(probe_ms), None)
    geom = (C.c_int32 * 4)()
    lib.fc_clash_geometry(N_ATOMS, geom)

    ap_kernel_ms = ap_ms.value / max(1, ap_n.value)
    ap_tf = FLOP_PER_PAIR * pairs / (ap_kernel_ms * 1e-3) / 1e12
    roofline_allpairs = {
        "bound": "fp32", "kernel": "fc::clash_f32_kernel", "achieved": ap_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": ap_tf / peak_tf, "traffic": None, "kernel_ms": ap_kernel_ms, "ms_per_step": ap_ms_per_step,
        "poses_per_s": world * n_poses / (ap_ms_per_step * 1e-3),
        "executed_tflops": EXEC_FLOP_PER_PAIR * pairs / (ap_kernel_ms * 1e-3) / 1e12,
        "executed_frac": EXEC_FLOP_PER_PAIR * pairs / (ap_kernel_ms * 1e-3) / 1e12 / peak_tf,
        "peak_source": f"{sms} SMs x 128 FP32 lanes x 2 x sm_max_mhz ({peak_kind} MEASURED_PEAKS.json clock); FFMA2 probe "
                       f"measured {probe_tf.value:.1f} TFLOP/s in this run",
        "note": "every atom pair evaluated (Gram form: 3 FMA + half an FMNMX3 per pair = 4 issue cycles per pair per "
                "lane, so frac = 1.0 is this formulation's ceiling and the FMA pipe cannot exceed 75 %); run with "
                "FC_CLASH_MODE=0 on the same poses; survivor bitmask identical to the cell-list screen's (asserted)"}

    # ---- roofline of the kernel that is timed.  The cell-list screen is bounded from below by its HBM traffic only
    #      (96 B of transform read + 1 bit written per pose); what actually limits it is instruction issue, so the
    #      issue-slot utilisation is reported beside it (warp instructions per pose from the committed ncu capture).
    pose_bytes = 96.0 + 0.125
    alg_bytes = n_poses * pose_bytes
    hbm_achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    prof, prof_src = _profile_json("r02_clash_cell_profile.json")
    traffic = None
    issue = None
    if prof:
        traffic = (prof["dram_read_bytes"] + prof["dram_write_bytes"]) / prof["poses"] * n_poses
        inst = prof["warp_inst_executed"] / prof["poses"] * n_poses
        issue_peak = sms * 4 * clock_hz
        issue = {"warp_inst_per_pose": prof["warp_inst_executed"] / prof["poses"], "achieved_inst_per_s": inst / (kernel_ms * 1e-3),
                 "peak_inst_per_s": issue_peak, "frac": inst / (kernel_ms * 1e-3) / issue_peak,
                 "source": f"{prof_src}: smsp__inst_executed.sum of the screen's level kernels over {prof['poses']} poses; peak = "
                           f"{sms} SMs x 4 schedulers x sm_max_mhz"}
    roofline = {
        "bound": "hbm", "kernel": "fc::clash_cell_kernel (3 level launches)" if cell_path else "fc::clash_f32_kernel",
        "achieved": hbm_achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": hbm_achieved / peaks.get("hbm_gbs", 6548.2),
        "traffic": traffic, "traffic_source": prof_src, "algorithmic_bytes": alg_bytes,
        "algorithmic_bytes_per_pose": pose_bytes, "peak_source": f"{peak_kind} MEASURED_PEAKS.json hbm_gbs (burst copy)",
        "kernel_ms": kernel_ms, "issue": issue,
        "work_avoided": {"allpairs_flop": FLOP_PER_PAIR * pairs, "allpairs_equivalent_tflops": FLOP_PER_PAIR * pairs / (kernel_ms * 1e-3) / 1e12,
                         "speedup_over_allpairs_kernel": ap_kernel_ms / kernel_ms,
                         "note": "SURVEY.md 8d counts 8 FLOP x all N_A x N_B atom pairs per pose; the cell-list screen reaches the "
                                 "same decisions (bitmask equality asserted on the full set) while visiting only pairs inside a "
                                 "candidate radius, so that figure measures work avoided and is NOT a roofline fraction"},
        "note": "HBM is the only hard floor of this kernel (it reads each pose once); it is instruction-issue bound "
                "(see `issue` and profiles/r02_clash_cell_ncu_summary.md), so frac is small by construction"}

    cpu = None
    if world == 1 and not args.no_cpu:
        cores, logical = physical_cores()
        cs = CpuScreen(cores)
        n_sample = 6000 * cores
        rate, _, wall = cs.rate(n_sample)
        cs.close()
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "logical_cpus": logical, "kind": "port",
               "sample": f"{n_sample} poses of the same sweep ({wall:.1f} s wall): per-pose get_embed + "
                         "scipy cdist compenetration_check (oracle/port.py), one worker per physical core, pool "
                         "created outside the timed window"}
        if os.path.isdir("/root/reference/firecode"):
            cs = CpuScreen(cores, use_reference=True)
            rate_ref, _, wall_ref = cs.rate(n_sample)
            cs.close()
            cpu["reference_function"] = {"value": rate_ref, "kind": cs.kind, "seconds": wall_ref,
                                         "note": "the unmodified firecode.utils.compenetration_check beside the port"}

    extras = None
    if world == 1 and not args.no_extras:
        extras = run_extras()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (+f64 recheck)", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "poses_per_gpu": n_poses, "n_atoms": [N_ATOMS, N_ATOMS], "l2": "inputs (960 MB of pose transforms per step) exceed L2",
                   "pass_fraction": n_pass / n_poses, "fp64_rechecks": n_recheck,
                   "path": "cell-list screen (default)" if cell_path else "all-pairs kernel (FC_CLASH_MODE=0)",
                   "checks": "survivor bitmasks of the cell-list screen, the all-pairs kernel, the compact-pose screen and "
                             "the host API are bit-identical on the full pose set (asserted in this run)",
                   "kernel_geometry": {"atoms_per_thread": geom[0], "threads_per_pose": geom[1],
                                       "poses_per_tile": geom[2], "threads_per_block": geom[3]}},
        "roofline": roofline,
        "roofline_allpairs": roofline_allpairs,
        "device_resident_pose7": {"value": world * n_poses / (q7_ms * 1e-3), "unit": UNIT, "ms_per_step": q7_ms,
                                  "note": "same step on the compact poses (28 B per pose read instead of 96 B)"},
        "strong_scaling": strong,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "clocks": clocks,
        "gpu_launches": launches_per_step * args.steps,
        "other_workloads": extras,
    }
    # second half of BASELINE.json's metric ("... + RMSD pairs/s"): the C4 pruning run, surfaced at the top level
    c4 = (extras or {}).get("C4_rmsd_pruning_200k")
    if c4_sharded:
        line["rmsd_pairs"] = c4_sharded
    if c4:
        line["rmsd_pairs"] = {"metric": "RMSD pairs/s", "value": c4["rmsd_pairs_per_s"], "unit": "pairs/s",
                              "workload": "C4: prune_by_rmsd of 200 k conformers x 120 atoms through the host API",
                              "seconds": c4["seconds"], "roofline": c4.get("roofline")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras():
    """Wall-clock throughput of the other BASELINE.json configs through the public host API
    (host buffers in, results out), one GPU.  Secondary numbers: the headline is the C3 sweep."""
    from firecode_b200 import embeds, problem, pruner, synthetic, torsion
    from firecode_b200.synthetic_embedder import make_embedder
    import networkx as nx

    out = {}

    def timed(fn, reps=2):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        return (time.perf_counter() - t0) / reps, r

    # C1: string embed, 2 x (10 conformers, 30 atoms), 2 centres, 36 angles = 14 400 tuples -- through the
    # reference-named entry point string_embed(embedder) (embeds.py:51-158): problem extraction + screen + coordinates
    emb = make_embedder("string", 10, 30, seed=synthetic.SEED, n_orb=2)
    dt, poses = timed(lambda: embeds.string_embed(emb), reps=5)
    rep = emb.b200_report
    out["C1_string_embed"] = {"api": "firecode_b200.embeds.string_embed(embedder)", "poses_per_s": rep.n_poses / dt,
                              "poses": rep.n_poses, "kept": rep.n_kept, "clash_pass": rep.n_clash_pass, "seconds": dt,
                              "returned_bytes": int(poses.nbytes)}
    # C2-like bimolecular cyclical embed: 2 x (50 conformers, 60 atoms), 4 pivots each, 36 angle pairs
    emb = make_embedder("cyclical", 50, 60, seed=synthetic.SEED + 1, n_reactive=2, n_orb=2)
    dt, poses = timed(lambda: embeds.cyclical_embed(emb), reps=2)
    rep = emb.b200_report
    out["C2_cyclical_embed_bimolecular"] = {"api": "firecode_b200.embeds.cyclical_embed(embedder)", "poses_per_s": rep.n_poses / dt,
                                            "poses": rep.n_poses, "kept": rep.n_kept, "clash_pass": rep.n_clash_pass,
                                            "seconds": dt, "returned_bytes": int(poses.nbytes)}
    # C2 (BASELINE.json configs[1]): trimolecular cyclical embed, 3 x (50 conformers, 60 atoms), one pivot per
    # molecule, 8 orientations, 216 angle triples = 125 000 conformer triples -> 1 M groups -> 216 M poses.
    # (a) the screen alone (kept indices, no coordinates), (b) the reference-named entry point, which also returns the
    # coordinates of every kept pose (float64, (kept, 180, 3)) as the reference's cyclical_embed does
    emb = make_embedder("cyclical", 50, 60, seed=synthetic.SEED + 2, n_mols=3, n_reactive=2, n_orb=1)
    tprob = problem.cyclical_problem(emb)
    dt, (poses, cons, rep) = timed(lambda: embeds.cyclical3_screen(tprob, want_status=False, want_coords=False), reps=1)
    out["C2_cyclical_embed_trimolecular"] = {
        "api": "firecode_b200.embeds.cyclical3_screen(problem, want_status=False, want_coords=False)",
        "poses_per_s": rep.n_poses / dt, "poses": rep.n_poses, "groups": int(len(rep.group_choice)),
        "kept": rep.n_kept, "clash_pass": rep.n_clash_pass, "seconds": dt,
        "min_direction_search_gap_deg": float(rep.group_gap.min()) if len(rep.group_gap) else None,
        "note": "full C2: group enumeration (C++), stateful 343-point direction search, 3 block screens over 36 "
                "distinct angle pairs each, keep-first RMSD; kept-pose coordinates not materialised"}
    t0 = time.perf_counter()
    poses = embeds.cyclical_embed(emb)
    dt = time.perf_counter() - t0
    rep = emb.b200_report
    out["C2_cyclical_embed_trimolecular_api"] = {
        "api": "firecode_b200.embeds.cyclical_embed(embedder)", "poses_per_s": rep.n_poses / dt, "poses": rep.n_poses,
        "kept": rep.n_kept, "seconds": dt, "returned_bytes": int(poses.nbytes),
        "note": "the drop-in call: same screen plus the float64 coordinates of every kept pose, streamed to the caller's array"}
    del poses
    # C4-like RMSD pruning: 20 000 conformers of a 120-atom molecule (400 basins)
    rng = np.random.default_rng(synthetic.SEED + 4)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 20000, 120, 400)
    dt, (kept, mask) = timed(lambda: pruner.prune_by_rmsd(structures, atoms, 0.5), reps=1)
    rep = pruner.last_report
    out["C4_rmsd_pruning_20k"] = {"rmsd_pairs_per_s": rep.pairs_tiled / dt, "pairs": rep.pairs_tiled,
                                  "pairs_eigen_solved": rep.pairs_solved, "passes": rep.passes,
                                  "kept": int(mask.sum()), "n": len(mask), "seconds": dt,
                                  "conventions": {"keep": rep.keep, "pass_mode": rep.pass_mode}}
    # C4 (BASELINE.json configs[3]) at full size: 200 000 conformers of a 120-atom molecule (2 000 basins)
    rng = np.random.default_rng(synthetic.SEED + 4)
    atoms, structures, _ = synthetic.pruning_ensemble(rng, 200000, 120, 2000)
    dt, (kept, mask) = timed(lambda: pruner.prune_by_rmsd(structures, atoms, 0.5), reps=1)
    rep = pruner.last_report
    out["C4_rmsd_pruning_200k"] = {"rmsd_pairs_per_s": rep.pairs_tiled / dt, "pairs": rep.pairs_tiled,
                                   "pairs_skipped_known_dissimilar": rep.pairs_skipped,
                                   "pairs_eigen_solved": rep.pairs_solved, "passes": rep.passes,
                                   "kept": int(mask.sum()), "n": len(mask), "seconds": dt,
                                   "mask_checksum": int(np.flatnonzero(mask).sum()),
                                   "h2d_bytes": int(structures.nbytes),
                                   "library_ms": rep.wall_ms,
                                   "screen_tiles_planned": rep.screen_tiles_planned,
                                   "screen_tiles_multiplied": rep.screen_tiles_multiplied,
                                   "screen_note": "a 128 x 16 tile whose structures' singular-value ranges are further apart than "
                                                  "the threshold (E >= sum (sigma_k(P) - sigma_k(Q))^2) is ruled out without "
                                                  "being multiplied; `pairs` counts every pair the screen decided",
                                   "conventions": {"keep": rep.keep, "pass_mode": rep.pass_mode}}
    if rep.screen_ms > 0:
        # roofline of the tensor-core screen (fc::gram_tc_kernel), summed over the launches of the run: algorithmic work per
        # pair = 51 N_h FLOP (covariance + norms + rotate-and-deviate); the tensor cores execute the covariance
        # only, 18 FLOP per atom slot (K padded to a multiple of 8) per pair slot of every 128 x 16 tile
        peaks, peak_kind = _peaks()
        # operands are FP16 (same tensor-core rate as bf16): the denominator is the MEASURED dense bf16 figure
        tf32_peak = float(peaks["bf16_tflops"])
        tf32_src = f"{peak_kind} MEASURED_PEAKS.json bf16_tflops (burst; FP16 operands run at the bf16 rate)"
        sec = rep.screen_ms * 1e-3
        kpad = 16 * ((rep.n_sel + 15) // 16)
        alg_tf = 51.0 * rep.n_sel * rep.pairs_tiled / sec / 1e12
        exe_tf = 18.0 * kpad * rep.screen_pair_slots / sec / 1e12
        out["C4_rmsd_pruning_200k"]["roofline"] = {
            "bound": "tensor", "kernel": "fc::gram_tc_kernel", "achieved": exe_tf, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": exe_tf / tf32_peak, "traffic": None, "algorithmic_tflops": alg_tf,
            "kernel_ms_total": rep.screen_ms, "launches": rep.screen_launches, "pairs": rep.pairs_tiled,
            "pair_slots": rep.screen_pair_slots, "candidates_to_fp64": rep.screen_candidates, "atoms_in_rmsd": rep.n_sel,
            "frac_note": "frac = EXECUTED tensor FLOP / measured dense 16-bit peak (what the tensor pipe did); the algorithmic figure of "
                         "SURVEY.md 8d (51 N_h FLOP per pair, never executed as such) is kept as algorithmic_tflops",
            "peak_source": tf32_src,
            "note": "FP16 Gram matrix of the centred heavy-atom coordinates (tcgen05.mma kind::f16 M128 N48 K16, FP32 accumulators in "
                    "TMEM); the kernel is bound by its FP32 epilogue (3x3 singular-value bounds of 2048 pairs per tile: 410 "
                    "warp instructions per epilogue warp and tile, profiles/r02_gram_tc_ncu_summary.md), not by tensor "
                    "throughput; pairs the screen cannot rule out are re-evaluated in FP64",
            "screen_pairs_per_s": rep.pairs_tiled / sec}
    del structures, kept, mask
    # TFD ensemble pruning (torsion_module.py:957-1043): 20 000 conformers of a 40-atom molecule, 12 quadruplets
    rng = np.random.default_rng(synthetic.SEED + 6)
    atoms, tstruct, _ = synthetic.pruning_ensemble(rng, 20000, 40, 2000, jitter=(0.0, 0.08))
    quads = np.array([rng.choice(40, 4, replace=False) for _ in range(12)], dtype=np.int64)
    dt, (tkept, tmask) = timed(lambda: torsion.prune_conformers_tfd(tstruct, quads), reps=1)
    out["TFD_pruning_20k"] = {"structures_per_s": len(tmask) / dt, "n": len(tmask), "kept": int(tmask.sum()),
                              "seconds": dt, "note": "fingerprints + first-match searches on the GPU, cluster "
                                                     "resolution (networkx, as the reference) on the host"}
    # C5: torsion scan, 1 000 conformers x 8 torsions x 36 steps of a 120-atom molecule
    rng = np.random.default_rng(synthetic.SEED + 5)
    atoms, coords, bonds, picks = synthetic.conformer_ensemble(rng, 1000, 120, n_torsions=8)
    g = nx.Graph(); g.add_nodes_from(range(120)); g.add_edges_from(bonds)
    tors = []
    for p, ch in picks:
        nb_p = [k for k in g.neighbors(p) if k != ch]; nb_c = [k for k in g.neighbors(ch) if k != p]
        if nb_p and nb_c:
            tors.append((nb_p[0], p, ch, nb_c[0]))
    masks = [torsion.get_rotation_mask(g, t) for t in tors]
    angles = np.arange(36) * 10.0
    dt, res = timed(lambda: torsion.torsion_scan(coords, tors, masks, angles, thresh=1.5, want_coords=False))
    n_items = len(coords) * len(tors) * len(angles)
    out["C5_torsion_scan"] = {"structures_per_s": n_items / dt, "structures": n_items,
                              "pass_fraction": float(res["passed"].mean()), "seconds": dt,
                              "note": "mask only; coordinates of 288 k x 120 atoms are optional output"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--poses", type=int, default=N_POSES, help="poses per GPU per step")
    ap.add_argument("--e2e-poses", type=int, default=N_POSES)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (C1, C2, C4, C5)")
    args = ap.parse_args()
    # the driver parses ONE JSON line from stdout: libraries that write to fd 1 (NCCL prints its version
    # banner there) are diverted to stderr, and only the result line goes to the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
