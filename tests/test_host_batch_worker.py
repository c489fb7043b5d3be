"""SURVEY.md §8 rows f1 + f2 in C++: host/packer.hpp and host/executor.hpp around the C ABI (tests/cpp/converge_test.cpp).
CPU tier links the hostsim test build, GPU tier the real libgarecon.so; both check every change set against the oracle."""
import json
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def _run(tmp_path, lib: Path, n: int, prog: str = "converge_test"):
    import __graft_entry__ as ge
    oracle = ge.build_oracle()
    exe = tmp_path / prog
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(REPO), "-o", str(exe), str(REPO / "tests" / "cpp" / f"{prog}.cpp"),
                    str(lib), str(oracle), f"-Wl,-rpath,{lib.parent}", f"-Wl,-rpath,{Path(oracle).parent}", "-pthread"], check=True)
    out = subprocess.run([str(exe), str(n)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def _check(rows, n):
    assert all(r["oracle_equal"] for r in rows)                       # every change set bit-exact vs the oracle
    by = {}
    for r in rows:
        by.setdefault(r["phase"], []).append(r)
    for phase, rs in by.items():
        assert rs[-1]["n_ops"] == 0, (phase, rs)                       # each phase reaches a fixed point
        assert len(rs) <= 4
    create, repair, cleanup = by["create"], by["repair"], by["cleanup"]
    assert create[0]["n_ops"] > n // 2 and create[0]["accelerators"] > n // 2      # round 0 creates the accelerators
    assert create[1]["records"] > create[0]["records"]                               # records follow once the accelerators exist
    assert create[-1]["not_ok_ga"] > 0                                               # provisioning LBs keep requeueing (30 s)
    # after the repair every object syncs cleanly, except the few that ask for records without managing an accelerator:
    # ListGlobalAcceleratorByHostname finds none, the reference requeues them every minute forever (route53.go:73-77)
    assert repair[0]["n_ops"] > 0 and repair[-1]["not_ok_ga"] == 0 and 0 < repair[-1]["not_ok_r53"] < n // 10
    assert cleanup[0]["n_ops"] > 0
    assert cleanup[-1]["accelerators"] < repair[-1]["accelerators"] and cleanup[-1]["records"] < repair[-1]["records"]


def test_batch_worker_converges_on_hostsim(tmp_path):
    import __graft_entry__ as ge
    rows = _run(tmp_path, Path(ge.build_hostsim()), 150)
    _check(rows, 150)


@pytest.mark.gpu
def test_batch_worker_converges_on_gpu(tmp_path):
    rows = _run(tmp_path, REPO / "aws-global-accelerator-controller_b200" / "libgarecon.so", 5000)
    _check(rows, 5000)


def test_executor_failure_semantics_on_hostsim(tmp_path):
    """Injected AWS errors: rollback of a partial create, the Ingress path's swallowed listener error, first error ends the
    object, events, rate-limited requeue of failed keys, per-controller execution, GAR_PENDING arguments, paginated packing,
    convergence through failures (tests/cpp/executor_faults_test.cpp holds the hand-derived expectations)."""
    import __graft_entry__ as ge
    rows = _run(tmp_path, Path(ge.build_hostsim()), 0, prog="executor_faults_test")
    assert rows[-1] == {"failed_checks": 0}


@pytest.mark.gpu
def test_executor_failure_semantics_on_gpu(tmp_path):
    rows = _run(tmp_path, REPO / "aws-global-accelerator-controller_b200" / "libgarecon.so", 0, prog="executor_faults_test")
    assert rows[-1] == {"failed_checks": 0}
