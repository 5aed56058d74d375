"""C1 (BASELINE.json configs[0]): string embed of two 10-conformer ensembles of 30 atoms, 14 400 tuples."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from firecode_b200 import embeds, problem, synthetic
from firecode_b200.synthetic_embedder import make_embedder
emb = make_embedder("string", 10, 30, seed=synthetic.SEED, n_orb=2)
prob = problem.string_problem(emb)
for i in range(4):
    t = time.perf_counter(); poses, rep = embeds.string_screen(prob); dt = time.perf_counter() - t
    print(f"C1: {rep.n_poses} tuples, pass {rep.n_clash_pass}, kept {rep.n_kept}, {dt*1e3:.2f} ms", flush=True)
for i in range(3):
    t = time.perf_counter(); prob = problem.string_problem(emb); t1 = time.perf_counter(); poses = embeds.string_embed(emb); dt = time.perf_counter() - t
    print(f"C1 via string_embed(embedder): {dt*1e3:.2f} ms (problem extraction alone {1e3*(t1-t):.2f} ms)", flush=True)
