"""One-off bug hunt: many more seeds than CI through hostsim (the device logic on the host) vs the oracle and pyref."""
import importlib, sys, copy, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / 'tests'))
import __graft_entry__ as ge
import randmodel, multilbi
garecon = importlib.import_module("aws-global-accelerator-controller_b200")
ob = importlib.import_module("oracle.binding")
pyref = importlib.import_module("oracle.pyref")
shard = importlib.import_module("aws-global-accelerator-controller_b200.shard")
lib = garecon.abi.load_library(ge.build_hostsim())
lo, hi = int(sys.argv[1]), int(sys.argv[2])
e = garecon.Engine(cluster_name="default", lib=lib)
bad = 0
t0=time.time()
for seed in range(lo, hi):
    for maker, kw in ((randmodel.make, dict(n_objects=20 + seed % 50)), (multilbi.make, dict(n_objects=10 + seed % 30))):
        objects, actual = maker(seed, **kw)
        snap = garecon.pack(objects, actual)
        want = ob.diff(snap, "default", mode=1)
        e.load(snap)
        got = e.diff()
        if got.diff(want) != []:
            bad += 1; print("HOSTSIM MISMATCH", maker.__module__, seed, got.describe_first_mismatch(want), flush=True)
        w0 = ob.diff(snap, "default", mode=0)
        if w0.diff(want) != []:
            bad += 1; print("ORACLE MODE0/1 MISMATCH", maker.__module__, seed, flush=True)
        w2 = ob.diff(snap, "default", mode=2, threads=1 + seed % 3)
        if w2.diff(want) != []:
            bad += 1; print("ORACLE MODE2 MISMATCH", maker.__module__, seed, flush=True)
        res = pyref.diff(copy.deepcopy(objects), copy.deepcopy(actual), "default")
        if list(want.status_ga) != res["status_ga"] or list(want.status_r53) != res["status_r53"] or [tuple(int(x) for x in op) for op in want.ops.tolist()] != res["ops"]:
            bad += 1; print("PYREF MISMATCH", maker.__module__, seed, flush=True)
    if seed % 7 == 0:  # sharded on hostsim
        objects, actual = randmodel.make(seed, n_objects=30 + seed % 40)
        g = 2 + seed % 4
        slices = shard.slice_model(objects, actual, g)
        engines, keep, snaps = [], [], []
        for objs_r, act_r, _ in slices:
            x = garecon.Engine(cluster_name="default", lib=lib); s = garecon.pack(objs_r, act_r); snaps.append(s); x.load(s); engines.append(x)
        if seed % 2: shard.exchange_local_peers(engines, [s[2] for s in slices])
        else: shard.exchange_local(engines, [s[2] for s in slices], keep)
        parts = [x.diff() for x in engines]
        merged = shard.merge_changesets(parts, len(objects))
        want = ob.diff(garecon.pack(objects, actual), "default", mode=1)
        if merged["ops"].tolist() != want.ops.tolist() or merged["status_ga"].tolist() != want.status_ga.tolist() or merged["status_r53"].tolist() != want.status_r53.tolist():
            bad += 1; print("SHARDED MISMATCH", seed, g, flush=True)
        for x in engines: x.close()
print("done", lo, hi, "bad", bad, "sec", round(time.time()-t0,1), flush=True)
