"""Sharded mode (SURVEY.md §8 row e, BASELINE configs[3]): one cluster cut into n_ranks slices, rows re-homed by key hash in
two exchanges, every shard running the ordinary pipeline on its self-contained sub-snapshot.  The merged result must equal
the unsharded change set of the whole cluster bit for bit (oracle on the concatenated model).

CPU tier: the device code compiled for the host (tests/hostsim), all ranks' engines in one process; a 2-process gloo test
covers the torch.distributed data path (tests/test_ranks_gloo.py style)."""
import importlib

import numpy as np
import pytest

import hotkeys
import randmodel

shard = importlib.import_module("aws-global-accelerator-controller_b200.shard")


@pytest.fixture(scope="module")
def hostlib(garecon):
    import __graft_entry__ as ge
    return garecon.abi.load_library(ge.build_hostsim())


def run_sharded(garecon, lib, objects, actual, n_ranks, cluster="default"):
    slices = shard.slice_model(objects, actual, n_ranks)
    engines, snaps, keep = [], [], []
    for objs_r, act_r, _ in slices:
        e = garecon.Engine(cluster_name=cluster, lib=lib)
        snap = garecon.pack(objs_r, act_r)
        e.load(snap)
        engines.append(e)
        snaps.append(snap)
    shard.exchange_local(engines, [s[2] for s in slices], keep)
    parts = [e.diff() for e in engines]
    for e in engines:
        e.close()
    return shard.merge_changesets(parts, len(objects)), parts


def check(garecon, oracle, lib, objects, actual, n_ranks, cluster="default"):
    got, parts = run_sharded(garecon, lib, objects, actual, n_ranks, cluster)
    want = oracle.diff(garecon.pack(objects, actual), cluster, mode=1)
    assert np.array_equal(got["status_ga"], want.status_ga), first_diff(got["status_ga"], want.status_ga)
    assert np.array_equal(got["status_r53"], want.status_r53), first_diff(got["status_r53"], want.status_r53)
    assert np.array_equal(got["derived"], want.derived)
    assert got["section_begin"].tolist() == want.section_begin.tolist()
    assert got["ops"].tolist() == want.ops.tolist(), first_diff(got["ops"].tolist(), want.ops.tolist())
    return got, parts


def first_diff(a, b):
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return f"first difference at {i}: {x} vs {y}"
    return f"lengths {len(a)} vs {len(b)}"


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("seed", range(6))
def test_sharded_matches_unsharded(garecon, oracle, hostlib, seed, n_ranks):
    objects, actual = randmodel.make(seed, n_objects=60)
    check(garecon, oracle, hostlib, objects, actual, n_ranks)


@pytest.mark.parametrize("seed", range(6, 30))
def test_sharded_more_seeds(garecon, oracle, hostlib, seed):
    objects, actual = randmodel.make(seed, n_objects=40)
    check(garecon, oracle, hostlib, objects, actual, 2 + seed % 3)


def test_sharded_hot_keys(garecon, oracle, hostlib):
    """Many owners claim the same record names / hostnames: alias records replicate to several homes, directory shards see
    long duplicate chains."""
    objects, actual = hotkeys.make()
    got, parts = check(garecon, oracle, hostlib, objects, actual, 4)
    assert len(got["ops"]) > 200
    assert sum(p.n_objects for p in parts) == len(objects)


def test_sharded_other_cluster_name(garecon, oracle, hostlib):
    objects, actual = randmodel.make(3, n_objects=40, cluster="prod-1")
    check(garecon, oracle, hostlib, objects, actual, 3, cluster="prod-1")


def test_sharded_empty_and_lopsided(garecon, oracle, hostlib):
    check(garecon, oracle, hostlib, [], {}, 2)
    objects, actual = randmodel.make(11, n_objects=5)  # fewer objects than ranks: some slices are empty
    check(garecon, oracle, hostlib, objects, actual, 8)
    check(garecon, oracle, hostlib, objects, {}, 3)
    check(garecon, oracle, hostlib, [], actual, 3)


def test_shards_are_balanced_and_disjoint(garecon, oracle, hostlib):
    objects, actual = randmodel.make(5, n_objects=400)
    got, parts = check(garecon, oracle, hostlib, objects, actual, 4)
    sizes = [p.n_objects for p in parts]
    assert sum(sizes) == 400 and min(sizes) > 50
