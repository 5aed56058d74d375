for c in 262144 524288 1048576; do echo -n "chunk=$c: "; FC_CLASH_CHUNK=$c python bench.py --steps 5 --no-cpu --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['e2e']['ms_per_step_each'])"; done
python - <<'PY'
import time, sys
sys.path.insert(0,'.')
from firecode_b200 import embeds, problem, synthetic
from firecode_b200.synthetic_embedder import make_embedder
emb = make_embedder("cyclical", 50, 60, seed=synthetic.SEED + 1, n_reactive=2, n_orb=2)
cprob = problem.cyclical_problem(emb)
for i in range(2):
    t=time.perf_counter(); g=embeds.cyclical_groups(cprob); t1=time.perf_counter()-t
    t=time.perf_counter(); poses, cons, rep = embeds.cyclical_screen(cprob, groups=g); t2=time.perf_counter()-t
    print("bimolecular: groups", len(g["conf"]), "table %.3fs screen %.3fs poses %d kept %d" % (t1, t2, rep.n_poses, rep.n_kept))
PY
