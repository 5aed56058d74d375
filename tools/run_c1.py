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
# where one call spends its time on the host side: the library call, then the accessors that fill numpy arrays
import ctypes as C
from firecode_b200 import _lib
lib = _lib.load(require_device=True)
for i in range(2):
    t0 = time.perf_counter(); c, keep = embeds._string_problem_c(prob); handle = C.c_void_p()
    t1 = time.perf_counter(); _lib.check(lib.fc_string_screen(C.byref(c), C.byref(handle)), "fc_string_screen")
    t2 = time.perf_counter(); res = embeds._Result(lib, handle); k = res.kept_indices(); st = res.status(); ti = res.ties()
    t3 = time.perf_counter(); poses = res.kept_coords()
    t4 = time.perf_counter(); res.close()
    t5 = time.perf_counter()
    print(f"C1 host breakdown: marshal {1e3*(t1-t0):.2f} ms, fc_string_screen {1e3*(t2-t1):.2f} ms, indices/status/ties "
          f"{1e3*(t3-t2):.2f} ms, kept_coords ({poses.nbytes/1e6:.1f} MB) {1e3*(t4-t3):.2f} ms, free {1e3*(t5-t4):.2f} ms", flush=True)
os.environ["FC_STRING_TRACE"] = "1"
poses, rep = embeds.string_screen(prob)
