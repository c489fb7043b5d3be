"""One-off bug hunt 2: incremental mode and EndpointGroupBinding diffs, many seeds, hostsim vs oracle (+ pyref for bindings)."""
import importlib, sys, random, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / 'tests'))
import __graft_entry__ as ge
import randmodel, multilbi, egbcases
import test_incremental as ti
garecon = importlib.import_module("aws-global-accelerator-controller_b200")
ob = importlib.import_module("oracle.binding")
pyref = importlib.import_module("oracle.pyref")
lib = garecon.abi.load_library(ge.build_hostsim())
lo, hi = int(sys.argv[1]), int(sys.argv[2])
e = garecon.Engine(cluster_name="default", lib=lib)
bad=0; t0=time.time()
for seed in range(lo, hi):
    rng = random.Random(seed)
    maker = randmodel.make if seed % 3 else multilbi.make
    objects, actual = maker(seed, n_objects=15 + seed % 45)
    snap = garecon.pack(objects, actual)
    rows, deleted = ti._pick(objects, actual, rng)
    e.load(snap)
    try:
        inc = e.diff_keys(rows, deleted)
        full = e.diff()
        want = ob.diff_keys(snap, rows, deleted, mode=0)
        if inc.diff(want) != []:
            bad+=1; print("KEYS MISMATCH", seed, inc.describe_first_mismatch(want), flush=True)
        ti._check_against_full(inc, full, rows)
    except AssertionError as ex:
        bad+=1; print("KEYS vs FULL MISMATCH", seed, flush=True)
    o2, a2, bindings, known = egbcases.random_bindings(seed)
    s2 = garecon.pack(o2, a2)
    pb = garecon.pack_bindings(bindings, known)
    e.load(s2)
    got = e.bindings_diff(pb)
    want = ob.bindings_diff(s2, pb)
    if got.status_ga.tolist() != want.status_ga.tolist() or got.ops.tolist() != want.ops.tolist():
        bad+=1; print("BINDINGS MISMATCH", seed, flush=True)
    st, ops = pyref.bindings_diff(o2, a2, bindings, set(known))
    if st != want.status_ga.tolist() or [tuple(int(x) for x in op) for op in want.ops.tolist()] != ops:
        bad+=1; print("BINDINGS PYREF MISMATCH", seed, flush=True)
print("done", lo, hi, "bad", bad, "sec", round(time.time()-t0,1), flush=True)
