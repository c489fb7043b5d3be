"""Tiny driver for ncu: generate one synthetic snapshot, load it, run the diff a few times.
usage: python profiles/run_once.py [config] [objects] [diffs]"""
import importlib
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pkg = importlib.import_module("aws-global-accelerator-controller_b200")
synth = importlib.import_module("aws-global-accelerator-controller_b200.synth")
c = synth.preset(cfg, n)
c.layout = 1  # column-major slabs: the bench default
snap = synth.SynthSnapshot(c)
with pkg.Engine(cluster_name=snap.cluster, reprepare=True) as e:
    e.load(snap)
    for _ in range(reps):
        cs = e.diff_device()
    print("ops", int(cs.n_ops), "kernels ms", cs.ms_kernels, "launches", cs.kernel_launches)
