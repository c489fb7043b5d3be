"""Two independent CPU restatements (oracle/oracle.cpp, oracle/pyref.py) and the oracle's evaluation modes must
agree bit for bit on randomized snapshots.  This is the stand-in for running the Go reference (no toolchain)."""
import importlib

import numpy as np
import pytest

import randmodel

pyref = importlib.import_module("oracle.pyref")


def _pyref_as_arrays(garecon, snap, res):
    abi = garecon.abi
    ops = np.array(res["ops"], dtype=np.uint32).reshape(-1, 6)
    tok_code = np.array([t[0] for t in res["tok"]], dtype=np.uint8)
    return ops, tok_code


@pytest.mark.parametrize("seed", range(40))
def test_oracle_modes_and_pyref_agree(garecon, oracle, seed):
    objects, actual = randmodel.make(seed, n_objects=30)
    snap = garecon.pack(objects, actual)
    faithful = oracle.diff(snap, "default", mode=0)
    indexed = oracle.diff(snap, "default", mode=1)
    mt = oracle.diff(snap, "default", mode=1, threads=3)
    assert faithful.diff(indexed) == [], faithful.describe_first_mismatch(indexed)
    assert faithful.diff(mt) == [], faithful.describe_first_mismatch(mt)
    for threads in (1, 5):  # mode 2: the tuned host-core baseline of bench.py decides exactly what the literal modes decide
        tuned = oracle.diff(snap, "default", mode=2, threads=threads)
        assert faithful.diff(tuned) == [], faithful.describe_first_mismatch(tuned)

    import copy
    res = pyref.diff(copy.deepcopy(objects), copy.deepcopy(actual), "default")
    assert list(faithful.status_ga) == res["status_ga"]
    assert list(faithful.status_r53) == res["status_r53"]
    assert list(faithful.derived) == res["derived"]
    assert list(faithful.section_begin) == res["section_begin"]
    got = [tuple(int(x) for x in op) for op in faithful.ops.tolist()]
    assert got == res["ops"]
    # tokeniser: codes, and names/regions as strings
    for i, (code, name, region) in enumerate(res["tok"]):
        assert faithful.tok_code[i] == code
        if code <= 2:
            assert snap.obj_str(faithful.tok_name[i]).decode() == name
            assert snap.obj_str(faithful.tok_region[i]).decode() == region
    # derived ports
    for i, p in enumerate(res["dports"]):
        b, e = faithful.dport_begin[i], faithful.dport_begin[i + 1]
        assert list(faithful.dports[b:e]) == (p or [])


def test_random_models_cover_the_op_space(garecon, oracle):
    """The randomized generator must actually reach every op code and status the path can produce."""
    ops, sts, dets = set(), set(), set()
    for seed in range(40):
        objects, actual = randmodel.make(seed, n_objects=30)
        cs = oracle.diff(garecon.pack(objects, actual), "default", mode=1)
        ops |= set((cs.ops["head"] & 0xFF).tolist())
        for arr in (cs.status_ga, cs.status_r53):
            sts |= set((arr & 0xFF).tolist())
            dets |= set(((arr >> 8) & 0xFF).tolist())
    assert ops == set(range(1, 11))
    assert sts >= {0, 1, 2, 3, 4, 5, 7}
    assert dets >= set(range(0, 12)) - {2}  # every detail code except the rare internal-ALB parse error is hit


@pytest.mark.parametrize("seed", range(30))
def test_tuned_baseline_equals_the_literal_port_on_multi_lbingress_models(garecon, oracle, seed):
    import multilbi
    snap = garecon.pack(*multilbi.make(seed))
    want = oracle.diff(snap, "default", mode=0)
    got = oracle.diff(snap, "default", mode=2, threads=1 + seed % 4)
    assert got.diff(want) == [], got.describe_first_mismatch(want)


def test_tuned_baseline_on_hot_keys_and_odd_clusters(garecon, oracle):
    import hotkeys
    snap = garecon.pack(*hotkeys.make())
    assert oracle.diff(snap, "default", mode=2, threads=4).diff(oracle.diff(snap, "default", mode=1)) == []
    objects, actual = randmodel.make(3, n_objects=40, cluster="prod-1")
    snap = garecon.pack(objects, actual)
    assert oracle.diff(snap, "prod-1", mode=2, threads=2).diff(oracle.diff(snap, "prod-1", mode=0)) == []
    empty = garecon.pack([], {})
    assert oracle.diff(empty, "default", mode=2, threads=3).diff(oracle.diff(empty, "default", mode=0)) == []
    only_actual = garecon.pack([], actual)
    assert oracle.diff(only_actual, "prod-1", mode=2, threads=3).diff(oracle.diff(only_actual, "prod-1", mode=0)) == []
