"""CPU tier: the device row logic + pipeline (csrc/gar_rows.h, gar_pipeline.h) compiled for the host
(tests/hostsim, a test build) must produce the oracle's change set bit for bit.  The same comparison runs
against the real sm_100a library in test_gpu_parity.py."""
import ctypes as C

import pytest

import randmodel


@pytest.fixture(scope="module")
def hostsim(garecon):
    import __graft_entry__ as ge
    path = ge.build_hostsim()
    lib = garecon.abi.load_library(path)
    e = garecon.Engine(cluster_name="default", lib=lib)
    yield e
    e.close()


@pytest.mark.parametrize("seed", range(60))
def test_hostsim_matches_oracle(garecon, oracle, hostsim, seed):
    objects, actual = randmodel.make(seed, n_objects=40)
    snap = garecon.pack(objects, actual)
    hostsim.load(snap)
    got = hostsim.diff()
    want = oracle.diff(snap, "default", mode=1)
    assert got.diff(want) == [], got.describe_first_mismatch(want)


def test_hostsim_empty_snapshot(garecon, oracle, hostsim):
    snap = garecon.pack([], {})
    hostsim.load(snap)
    got = hostsim.diff()
    want = oracle.diff(snap, "default", mode=1)
    assert got.diff(want) == []
    assert len(got.ops) == 0


def test_hostsim_other_cluster_name(garecon, oracle):
    import __graft_entry__ as ge
    lib = garecon.abi.load_library(ge.build_hostsim())
    objects, actual = randmodel.make(3, n_objects=40, cluster="prod-1")
    snap = garecon.pack(objects, actual)
    with garecon.Engine(cluster_name="prod-1", lib=lib) as e:
        e.load(snap)
        got = e.diff()
    want = oracle.diff(snap, "prod-1", mode=1)
    assert got.diff(want) == [], got.describe_first_mismatch(want)
    assert len(got.ops) > 10
