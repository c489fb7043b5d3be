"""CPU tier: the synthetic BASELINE configs are valid snapshots, exercise every op, and the oracle's evaluation
modes agree on them (indexed == faithful at sizes the faithful mode finishes in seconds)."""
import importlib

import numpy as np
import pytest

synth = None


@pytest.fixture(scope="module", autouse=True)
def _load(garecon):
    global synth
    import __graft_entry__ as ge
    ge.build_synth()
    synth = importlib.import_module("aws-global-accelerator-controller_b200.synth")


@pytest.mark.parametrize("cfg,n", [(1, 100), (2, 3000), (3, 3000), (5, 3000)])
def test_faithful_equals_indexed_on_baseline_configs(oracle, cfg, n):
    s = synth.generate(cfg, n)
    faithful = oracle.diff(s, "default", mode=0)
    indexed = oracle.diff(s, "default", mode=1)
    mt = oracle.diff(s, "default", mode=1, threads=4)
    assert faithful.diff(indexed) == [], faithful.describe_first_mismatch(indexed)
    assert faithful.diff(mt) == []
    tuned = oracle.diff(s, "default", mode=2, threads=4)
    assert faithful.diff(tuned) == [], faithful.describe_first_mismatch(tuned)


@pytest.mark.parametrize("cfg,n,layout", [(2, 60000, 0), (3, 60000, 1), (4, 60000, 1), (5, 60000, 0)])
def test_tuned_baseline_equals_indexed_at_size(oracle, cfg, n, layout):
    """The tuned CPU baseline (bench.py cpu_baseline, --impl reference) against the literal indexed port on every BASELINE
    distribution, both slab layouts, several thread counts."""
    c = synth.preset(cfg, n)
    c.layout = layout
    s = synth.SynthSnapshot(c)
    want = oracle.diff(s, "default", mode=1, threads=4)
    for threads in (1, 7):
        got = oracle.diff(s, "default", mode=2, threads=threads)
        assert got.diff(want) == [], got.describe_first_mismatch(want)


def test_cfg1_is_the_reference_fixture_shape(oracle):
    """BASELINE config 1: 100 Services type LoadBalancer (local_e2e/pkg/fixtures/service.go:10-51 shape)."""
    s = synth.generate(1, 100)
    assert s.objects.n_objects == 100
    cs = oracle.diff(s, "default", mode=0)
    assert ((cs.status_ga & 0xFF) != 0).all()      # every Service passes the controller filter
    assert set((cs.ops["head"] & 0xFF).tolist()) >= {1, 8}  # creates happen for the missing ones


def test_configs_reach_the_op_space(oracle):
    s = synth.generate(3, 20000)
    cs = oracle.diff(s, "default", mode=1, threads=4)
    ops = set((cs.ops["head"] & 0xFF).tolist())
    assert ops == set(range(1, 11)), ops
    codes = set((cs.status_ga & 0xFF).tolist()) | set((cs.status_r53 & 0xFF).tolist())
    assert codes >= {0, 1, 3, 4}
    # sanity on the mix: most objects are in sync, so ops per object stays well below 2
    assert 0.2 < len(cs.ops) / s.objects.n_objects < 2.5


def test_generator_is_deterministic(oracle):
    a = oracle.diff(synth.generate(2, 2000), "default", mode=1)
    b = oracle.diff(synth.generate(2, 2000), "default", mode=1)
    assert a.checksum() == b.checksum()
    c = oracle.diff(synth.generate(2, 2000, seed=99), "default", mode=1)
    assert a.checksum() != c.checksum()
