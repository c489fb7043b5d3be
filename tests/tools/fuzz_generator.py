"""One-off bug hunt 3: the synthetic generator (all presets, both slab layouts, many seeds, small sizes) through hostsim vs oracle;
and generator slices through the sharded path."""
import importlib, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / 'tests'))
import __graft_entry__ as ge
garecon = importlib.import_module("aws-global-accelerator-controller_b200")
synth = importlib.import_module("aws-global-accelerator-controller_b200.synth")
shard = importlib.import_module("aws-global-accelerator-controller_b200.shard")
ob = importlib.import_module("oracle.binding")
lib = garecon.abi.load_library(ge.build_hostsim())
lo, hi = int(sys.argv[1]), int(sys.argv[2])
e = garecon.Engine(cluster_name="default", lib=lib)
bad=0; t0=time.time()
for seed in range(lo, hi):
    for cfg in (1, 2, 3, 4, 5):
        n = 200 + (seed * 37 + cfg * 101) % 1500
        snap = synth.generate(cfg, n, seed=seed, layout=seed % 2)
        want = ob.diff(snap, snap.cluster, mode=1, threads=2)
        e2 = garecon.Engine(cluster_name=snap.cluster, lib=lib)
        e2.load(snap)
        got = e2.diff()
        if got.diff(want) != []:
            bad+=1; print("SYNTH MISMATCH", cfg, n, seed, got.describe_first_mismatch(want), flush=True)
        w2 = ob.diff(snap, snap.cluster, mode=2, threads=3)
        if w2.diff(want) != []:
            bad+=1; print("TUNED MISMATCH", cfg, n, seed, flush=True)
        e2.close()
    if seed % 3 == 0:
        cfg = 3 + seed % 3; g = 2 + seed % 3; n = 600 + seed % 900; n -= n % g
        slices = synth.cluster_slices(cfg, n, g, seed=seed, layout=seed % 2)
        bases = garecon.tables.shard_bases(slices)
        engines, keep, snaps = [], [], []
        for o, a in slices:
            x = garecon.Engine(cluster_name="default", lib=lib); s = garecon.tables.from_columns(o, a); snaps.append(s); x.load(s); engines.append(x)
        if seed % 2: shard.exchange_local_peers(engines, bases)
        else: shard.exchange_local(engines, bases, keep)
        parts = [x.diff() for x in engines]
        union = garecon.tables.concat_slices(slices)
        want = ob.diff(union, "default", mode=1, threads=2)
        merged = shard.merge_changesets(parts, int(union.objects.n_objects))
        if merged["ops"].tolist() != want.ops.tolist() or merged["status_ga"].tolist() != want.status_ga.tolist() or merged["status_r53"].tolist() != want.status_r53.tolist():
            bad+=1; print("SYNTH SHARDED MISMATCH", cfg, n, g, seed, flush=True)
        for x in engines: x.close()
print("done", lo, hi, "bad", bad, "sec", round(time.time()-t0,1), flush=True)
