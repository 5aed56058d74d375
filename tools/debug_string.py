import sys, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from firecode_b200 import embeds, problem, _lib
from oracle import port
from synth_embedder import make_embedder
emb = make_embedder("string", 2, 12, seed=9, n_orb=1)
prob = problem.string_problem(emb)
lib = _lib.load(True)
c, keep = embeds._string_problem_c(prob)
h = C.c_void_p()
_lib.check(lib.fc_string_stage1(C.byref(c), 0, -1, C.byref(h)), "stage1")
res = embeds._Result(lib, h)
surv = res.survivors(); fps = res.fingerprints()
print("n_surv", len(surv), "quads", prob.quadruplets.tolist())
for s, fp in zip(surv, fps):
    c1, c2, a1, a2, ang = prob.decode(s)
    rot, pos = port.string_transform(prob, c1, c2, a1, a2, ang)
    st = np.concatenate([prob.coords[0][c1], (rot @ prob.coords[1][c2].T).T + pos])
    ref = port.torsion_fingerprint(st, prob.quadruplets)
    d = np.abs(ref - fp); d = np.minimum(d, 360 - d)
    print(s, "maxdiff", d.max(), np.round(fp, 3), np.round(ref, 3))
