"""GPU tier: BASELINE config 1 — 100 synthetic Services through the batch worker (C++ mirror of the reference's
ProcessNextWorkItem / reconcileHandler, host/reconcile.hpp) over the C ABI.  The queue actions must be exactly what
reconcile.go:70-90 prescribes for the status words the diff produced."""
import importlib
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "aws-global-accelerator-controller_b200"


def _build(tmp_path):
    exe = tmp_path / "batch_worker_test"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(REPO), "-o", str(exe), str(REPO / "tests" / "cpp" / "batch_worker_test.cpp"),
                    f"-L{PKG}", f"-L{PKG / 'synth'}", "-lgarecon", "-lgarecon_synth", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{PKG / 'synth'}", "-pthread"], check=True)
    return exe


@pytest.mark.parametrize("cfg,n", [(1, 100), (2, 5000)])
def test_batch_worker_queue_actions(garecon, engine, tmp_path, cfg, n):
    import __graft_entry__ as ge
    ge.ensure_built()
    exe = _build(tmp_path)
    out = subprocess.run([str(exe), str(cfg), str(n)], capture_output=True, text=True, check=True).stdout
    got = json.loads(out)
    synth = importlib.import_module("aws-global-accelerator-controller_b200.synth")
    snap = synth.generate(cfg, n)
    engine.load(snap)
    cs = engine.diff()
    kinds = np.ctypeslib.as_array(snap.objects.obj_kind, shape=(n,))
    st = cs.status_ga[kinds == 0] & 0xFF
    n_svc = int((kinds == 0).sum())
    exp_forget = int(np.isin(st, [0, 1, 2]).sum())
    exp_after30, exp_after60 = int((st == 3).sum()), int((st == 4).sum())
    exp_rl = int((st == 5).sum())
    assert got["rc"] == 0 and got["keys"] == n_svc + 1 and got["deleted_keys"] == 1
    assert got["q_done"] == n_svc + 1
    assert got["q_ratelimited"] == exp_rl
    assert got["q_after30"] == exp_after30 and got["q_after60"] == exp_after60
    # RequeueAfter does Forget + AddAfter (reconcile.go:79-82); the deleted key is forgotten too
    assert got["q_forget"] == exp_forget + exp_after30 + exp_after60 + 1
    assert got["n_ops"] == len(cs.ops)
