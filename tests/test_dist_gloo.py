"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, variable-length
all-gather, bitmask exchange and the ordered keep-first merge.  The per-shard compute is injected
(the oracle plays the kernels' role here); on the GPU box the same functions run over NCCL."""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from firecode_b200 import dist as fdist
        from firecode_b200 import problem, synthetic
        from oracle import port as oport
        from synth_embedder import make_embedder

        # 1) variable-length gather
        x = np.arange(rank * 10, rank * 10 + 3 + 2 * rank, dtype=np.float64).reshape(-1, 1) * np.ones((1, 4))
        g = fdist.all_gather_varlen(x)
        assert g.shape == (3 + 5, 4) and g[0, 0] == 0 and g[3, 0] == 10

        # 2) sharded clash screen == single-process oracle
        rng = np.random.default_rng(7)
        _, a, _, _ = synthetic.molecule_cloud(rng, 40)
        _, b, _, _ = synthetic.molecule_cloud(rng, 33)
        xf = synthetic.sweep_poses(rng, a, b, 1001)
        full = oport.clash_batch(a, b, xf, thresh=1.5)[0]
        mask = fdist.clash_screen_sharded(a, b, xf, 1.5, screen=lambda fa, fb, x: oport.clash_batch(fa, fb, x, thresh=1.5)[0])
        assert np.array_equal(mask, full)

        # 3) ordered keep-first merge of the string embed == single-process oracle
        emb = make_embedder("string", 2, 14, seed=3, n_orb=2)
        prob = problem.string_problem(emb)
        ref = oport.string_embed(prob, want_poses=False)
        lo, hi = fdist.shard_bounds(prob.n_poses, world, rank)
        labels = np.array([p for p in range(lo, hi) if ref["clash_pass"][p]], dtype=np.int64)
        fps = []
        for p in labels:
            c1, c2, a1, a2, ang = prob.decode(p)
            rot, pos = oport.string_transform(prob, c1, c2, a1, a2, ang)
            st = np.concatenate([prob.coords[0][c1], (rot @ prob.coords[1][c2].T).T + pos])
            fps.append(oport.torsion_fingerprint(st, prob.quadruplets))
        fps = np.array(fps).reshape(len(labels), len(prob.quadruplets))

        def keep_fn(fp, lab):
            keep, acc = np.zeros(len(lab), dtype=bool), []
            for i in range(len(lab)):
                if not any(oport.tfd_sum(fp[i], fp[j]) < 10.0 for j in acc):
                    acc.append(i)
                    keep[i] = True
            return keep

        kept, all_labels, _ = fdist.ordered_keep_first(labels, fps, keep_fn)
        assert np.array_equal(all_labels, np.flatnonzero(ref["clash_pass"]))
        assert np.array_equal(kept, ref["kept"])

        # 4) cyclical embeds sharded over whole groups == single-process oracle (tri- and bimolecular)
        emb3 = make_embedder("cyclical", [2, 1, 1], 10, seed=17, n_mols=3, n_reactive=2, n_orb=1)
        prob3 = problem.cyclical_problem(emb3)
        ref3 = oport.cyclical_embed_trimol(prob3)

        def screen3(pb, lo, hi):
            o = oport.cyclical_embed_trimol(pb, conf_tuple_range=(lo, hi))
            return o["poses"], o["constrained"], len(o["clash_pass"]), o["kept"]

        poses3 = fdist.cyclical_embed_sharded(emb3, screen=screen3)
        assert np.array_equal(poses3, ref3["poses"]) and np.array_equal(emb3.b200_kept_indices, ref3["kept"])
        assert np.array_equal(emb3.constrained_indices, ref3["constrained"])

        # 5) torsion scan sharded over conformers == single-process oracle
        import networkx as nx

        rng = np.random.default_rng(3)
        atoms, cc, bonds, picks = synthetic.conformer_ensemble(rng, 5, 24, n_torsions=3)
        g = nx.Graph(); g.add_nodes_from(range(24)); g.add_edges_from(bonds)
        tors = []
        for p_, ch in picks:
            nb_p = [k for k in g.neighbors(p_) if k != ch]; nb_c = [k for k in g.neighbors(ch) if k != p_]
            if nb_p and nb_c:
                tors.append((nb_p[0], p_, ch, nb_c[0]))
        tmasks = [oport.rotation_mask(g, t) for t in tors]
        angs = np.arange(6) * 60.0
        full = oport.torsion_scan(cc, tors, tmasks, angs)[1]
        got = fdist.torsion_scan_sharded(cc, tors, tmasks, angs, scan=lambda x: oport.torsion_scan(x, tors, tmasks, angs)[1])
        assert np.array_equal(got, full)
        q.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        import traceback

        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_shard_bounds_and_bits():
    from firecode_b200 import dist as fdist

    assert [fdist.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert fdist.shard_bounds(0, 2, 1) == (0, 0)
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 1000):
        m = rng.random(n) < 0.4
        w = fdist.pack_bits(m)
        assert len(w) == (n + 31) // 32 and np.array_equal(fdist.unpack_bits(w, n), m)
        assert all(((int(w[i >> 5]) >> (i & 31)) & 1) == int(m[i]) for i in range(n))
