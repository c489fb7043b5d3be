"""Driver for ncu / compute-sanitizer of the sharded mode on ONE GPU: n_ranks engines on cuda:0, blobs exchanged inside the
process (shard.exchange_local), then each shard's diff.  Optionally also an EndpointGroupBinding diff.
usage: python profiles/run_sharded_once.py [config] [objects_total] [n_ranks] [reps] [bindings]"""
import importlib
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
g = int(sys.argv[3]) if len(sys.argv) > 3 else 4
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
with_bindings = len(sys.argv) > 5 and sys.argv[5] == "bindings"
pkg = importlib.import_module("aws-global-accelerator-controller_b200")
synth = importlib.import_module("aws-global-accelerator-controller_b200.synth")
shard = importlib.import_module("aws-global-accelerator-controller_b200.shard")
slices = synth.cluster_slices(cfg, n, g, layout=1, threads=8)  # column-major slabs, as the packer writes them
bases = pkg.tables.shard_bases(slices)
engines = []
for o, a in slices:
    e = pkg.Engine(cluster_name="default")
    e.load(pkg.tables.from_columns(o, a))
    engines.append(e)
for _ in range(reps):
    keep = []
    shard.exchange_local(engines, bases, keep, device="cuda:0")
    tot, ck = 0, 0
    for e in engines:
        cs = e.diff()
        tot += len(cs.ops)
        ck ^= cs.checksum()
    print("sharded step: ops", tot, "checksum", hex(ck), "launches (rank 0)", cs.kernel_launches)
if with_bindings:
    import egbcases
    objects, actual, bindings, known = egbcases.random_bindings(3, n_objects=200, n_bindings=2000)
    with pkg.Engine(cluster_name="default") as e:
        e.load(pkg.pack(objects, actual))
        out = e.bindings_diff(pkg.pack_bindings(bindings, known))
        print("bindings ops", len(out.ops))
for e in engines:
    e.close()
