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


def test_hostsim_hot_keys_force_the_fallback_paths(garecon, oracle, hostsim):
    import hotkeys
    objects, actual = hotkeys.make()
    snap = garecon.pack(objects, actual)
    hostsim.load(snap)
    got = hostsim.diff()
    want = oracle.diff(snap, "default", mode=1)
    faithful = oracle.diff(snap, "default", mode=0)
    assert want.diff(faithful) == []
    assert got.diff(want) == [], got.describe_first_mismatch(want)
    assert len(got.ops) > 200


def test_warp_emulation_catches_a_non_uniform_vote(garecon):
    """The hostsim warp emulator must flag a GAR_ANY that not every lane reaches (it would hang a real GPU)."""
    import ctypes, subprocess, tempfile, textwrap
    from pathlib import Path
    repo = Path(__file__).resolve().parent.parent
    src = textwrap.dedent(r'''
        #define main hostsim_unused_main
        #include "tests/hostsim/hostsim.cpp"
        #undef main
        struct Bad { GAR_HD void operator()(u32 i, bool valid) const { if (i % 2 == 0) (void)GAR_ANY(true); } };
        struct Good { GAR_HD void operator()(u32 i, bool valid) const { for (u32 k = 0; GAR_ANY(k < i % 5); k++) {} } };
        extern "C" int probe(int which) {
          gar_engine e;
          g_nonuniform_vote = false;
          if (which) e.for_each_warp("bad", 64, Bad{}); else e.for_each_warp("good", 64, Good{});
          return g_nonuniform_vote ? 1 : 0;
        }
    ''')
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "t.cpp").write_text(src)
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", str(repo), "-o", f"{d}/t.so", f"{d}/t.cpp"], check=True)
        lib = ctypes.CDLL(f"{d}/t.so")
        assert lib.probe(0) == 0
        assert lib.probe(1) == 1


def test_hostsim_tokeniser_fuzz(garecon, oracle, hostsim):
    import fuzzcases
    snap = garecon.pack(fuzzcases.fuzz_hostnames(11, 1500), {})
    hostsim.load(snap)
    got = hostsim.diff()
    want = oracle.diff(snap, "default", mode=1)
    assert got.diff(want) == [], got.describe_first_mismatch(want)
    assert len(set(got.tok_code.tolist())) >= 8


def test_hostsim_listen_ports_fuzz(garecon, oracle, hostsim):
    import fuzzcases
    snap = garecon.pack(fuzzcases.listen_objects(), {})
    hostsim.load(snap)
    got = hostsim.diff()
    want = oracle.diff(snap, "default", mode=1)
    assert got.diff(want) == [], got.describe_first_mismatch(want)
    assert len(got.dports) > 10
