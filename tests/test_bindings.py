"""EndpointGroupBinding set-diff: gar_bindings_diff against the oracle's literal restatement of
pkg/controller/endpointgroupbinding/reconcile.go.  The reference has no unit test for this controller (its e2e suite
needs a live AWS account), so the expectations below are hand-derived from the Go source and say so: parity unpinned."""
import pytest

import egbcases

ST = lambda w: w & 0xFF
DETAIL = lambda w: (w >> 8) & 0xFF


def _ops_of(cs, k):
    return [(int(op[0]) & 0xFF, int(op[3])) for op in cs.ops.tolist() if int(op[1]) == k]


def test_oracle_hand_derived_expectations(garecon, oracle):
    """The branches of reconcile.go, read off the Go source by hand (row numbers: ep rows are global CSR rows)."""
    abi = garecon.abi
    objects, actual, bindings, known = egbcases.hand_cases()
    snap = garecon.pack(objects, actual)
    b = garecon.pack_bindings(bindings, known)
    cs = oracle.bindings_diff(snap, b)
    ep0 = b.arrays["egb_ep_begin"].tolist()
    st = cs.status_ga.tolist()
    RM, ADD, W, UPD = abi.OP_EGB_REMOVE_ENDPOINT, abi.OP_EGB_ADD_ENDPOINT, abi.OP_EGB_UPDATE_WEIGHT, abi.OP_EGB_UPDATE_STATUS
    NONE = 0xFFFFFFFF
    # reconcileDelete
    assert (ST(st[0]), _ops_of(cs, 0)) == (abi.ST_OK, [(abi.OP_EGB_REMOVE_FINALIZER, NONE)])
    assert (ST(st[1]), _ops_of(cs, 1)) == (abi.ST_OK, [(abi.OP_EGB_REMOVE_FINALIZER, NONE)])
    assert (ST(st[2]), _ops_of(cs, 2)) == (abi.ST_REQUEUE_1S, [(RM, ep0[2]), (UPD, NONE)])
    # n=2: iteration 0 removes ids[0] and shifts ids[1] down; iteration 1 reads the stale copy at [1], then slices [2:] of a len-1 slice
    assert (ST(st[3]), _ops_of(cs, 3)) == (abi.ST_PANIC, [(RM, ep0[3]), (RM, ep0[3] + 1)])
    # n=3: removes ids[0], ids[2] (shifted into [1]), then [2] again (stale), panics on [3:] of len 1
    assert (ST(st[4]), _ops_of(cs, 4)) == (abi.ST_PANIC, [(RM, ep0[4]), (RM, ep0[4] + 2), (RM, ep0[4] + 2)])
    assert ST(st[5]) == abi.ST_PANIC and [a for _, a in _ops_of(cs, 5)] == [ep0[5], ep0[5] + 2, ep0[5] + 4, ep0[5] + 4]
    assert (ST(st[6]), _ops_of(cs, 6)) == (abi.ST_OK, [(abi.OP_EGB_REMOVE_FINALIZER, NONE)])
    # reconcileCreate never looks at the reference
    assert _ops_of(cs, 7) == [(abi.OP_EGB_ADD_FINALIZER, NONE)] and _ops_of(cs, 8) == [(abi.OP_EGB_ADD_FINALIZER, NONE)]
    # reconcileUpdate
    assert (ST(st[9]), _ops_of(cs, 9)) == (abi.ST_OK, [])
    assert (ST(st[10]), _ops_of(cs, 10)) == (abi.ST_OK, [(W, 0), (UPD, NONE)])
    assert _ops_of(cs, 11) == [(ADD, 0), (W, 0), (UPD, NONE)]
    assert _ops_of(cs, 12) == [(ADD, 0), (ADD, 1), (W, 0), (W, 1), (UPD, NONE)]
    assert _ops_of(cs, 13) == [(ADD, 0), (W, 0), (W, 1), (UPD, NONE)]
    assert _ops_of(cs, 14) == [(RM, ep0[14]), (RM, ep0[14] + 2), (ADD, 0), (W, 0), (W, 1), (UPD, NONE)]
    assert _ops_of(cs, 15) == [(ADD, 1), (ADD, 2), (W, 1), (W, 2), (UPD, NONE)]
    assert _ops_of(cs, 16) == [(RM, ep0[16]), (RM, ep0[16] + 1), (ADD, 1), (ADD, 2), (W, 1), (W, 2), (UPD, NONE)]
    assert (ST(st[17]), _ops_of(cs, 17)) == (abi.ST_OK, [])
    assert (ST(st[18]), _ops_of(cs, 18)) == (abi.ST_OK, [(UPD, NONE)])
    assert (ST(st[19]), _ops_of(cs, 19)) == (abi.ST_PANIC, [])
    assert (ST(st[20]), _ops_of(cs, 20)) == (abi.ST_PANIC, [])
    assert (ST(st[21]), _ops_of(cs, 21)) == (abi.ST_OK, [])
    assert (ST(st[22]), DETAIL(st[22])) == (abi.ST_ERR_RETRY, abi.D_REF_NOT_FOUND)
    assert (ST(st[23]), DETAIL(st[23])) == (abi.ST_ERR_RETRY, abi.D_REF_NOT_FOUND)
    assert _ops_of(cs, 24) == [(ADD, 6), (W, 6), (UPD, NONE)]
    assert _ops_of(cs, 25) == [(ADD, 2), (W, 2), (UPD, NONE)]
    assert (ST(st[26]), DETAIL(st[26])) == (abi.ST_ERR_RETRY, abi.D_REF_NOT_FOUND)
    assert (ST(st[27]), DETAIL(st[27])) == (abi.ST_ERR_RETRY, abi.D_NOT_ELB)
    assert (ST(st[28]), DETAIL(st[28])) == (abi.ST_ERR_RETRY, abi.D_LB_NOT_FOUND)
    assert (ST(st[29]), _ops_of(cs, 29)) == (abi.ST_REQUEUE_30S, [])
    assert (ST(st[30]), _ops_of(cs, 30)) == (abi.ST_REQUEUE_30S, [(RM, ep0[30])])
    assert (ST(st[31]), DETAIL(st[31]), _ops_of(cs, 31)) == (abi.ST_ERR_RETRY, abi.D_EG_NOT_FOUND, [])
    assert (ST(st[32]), _ops_of(cs, 32)) == (abi.ST_OK, [])
    assert (ST(st[33]), DETAIL(st[33]), _ops_of(cs, 33)) == (abi.ST_ERR_RETRY, abi.D_LB_NOT_FOUND, [])
    assert (ST(st[34]), DETAIL(st[34]), _ops_of(cs, 34)) == (abi.ST_ERR_RETRY, abi.D_LB_NOT_FOUND, [])
    assert _ops_of(cs, 35) == [(ADD, 4), (W, 0), (W, 4), (UPD, NONE)]
    assert _ops_of(cs, 36) == [(RM, ep0[36]), (ADD, 0), (W, 0), (UPD, NONE)]


@pytest.fixture(scope="module")
def hostsim(garecon):
    import __graft_entry__ as ge
    lib = garecon.abi.load_library(ge.build_hostsim())
    e = garecon.Engine(cluster_name="default", lib=lib)
    yield e
    e.close()


def _check(garecon, oracle, engine, objects, actual, bindings, known):
    snap = garecon.pack(objects, actual)
    b = garecon.pack_bindings(bindings, known)
    engine.load(snap)
    got = engine.bindings_diff(b)
    want = oracle.bindings_diff(snap, b)
    assert got.status_ga.tolist() == want.status_ga.tolist()
    assert got.ops.tolist() == want.ops.tolist()
    return got


def test_hostsim_hand_cases(garecon, oracle, hostsim):
    got = _check(garecon, oracle, hostsim, *egbcases.hand_cases())
    assert len(got.ops) > 40


@pytest.mark.parametrize("seed", range(20))
def test_hostsim_random_bindings(garecon, oracle, hostsim, seed):
    _check(garecon, oracle, hostsim, *egbcases.random_bindings(seed))


def test_hostsim_bindings_then_diff_share_one_snapshot(garecon, oracle, hostsim):
    """gar_bindings_diff must leave the loaded snapshot and its prepared state usable for gar_diff."""
    objects, actual, bindings, known = egbcases.random_bindings(3)
    snap = garecon.pack(objects, actual)
    b = garecon.pack_bindings(bindings, known)
    hostsim.load(snap)
    first = hostsim.diff()
    got = hostsim.bindings_diff(b)
    again = hostsim.diff()
    assert first.diff(again) == []
    assert got.ops.tolist() == oracle.bindings_diff(snap, b).ops.tolist()
    assert again.diff(oracle.diff(snap, "default", mode=1)) == []


def test_hostsim_no_bindings(garecon, oracle, hostsim):
    objects, actual, _, known = egbcases.random_bindings(1)
    got = _check(garecon, oracle, hostsim, objects, actual, [], known)
    assert got.n_objects == 0 and len(got.ops) == 0


@pytest.mark.gpu
def test_gpu_hand_cases(garecon, oracle, engine):
    got = _check(garecon, oracle, engine, *egbcases.hand_cases())
    assert got.kernel_launches > 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_gpu_random_bindings(garecon, oracle, engine, seed):
    _check(garecon, oracle, engine, *egbcases.random_bindings(seed, n_objects=80, n_bindings=400))


@pytest.mark.gpu
def test_gpu_bindings_then_diff_share_one_snapshot(garecon, oracle, engine):
    objects, actual, bindings, known = egbcases.random_bindings(5, n_objects=80, n_bindings=300)
    snap = garecon.pack(objects, actual)
    b = garecon.pack_bindings(bindings, known)
    engine.load(snap)
    got = engine.bindings_diff(b)
    again = engine.diff()
    assert got.ops.tolist() == oracle.bindings_diff(snap, b).ops.tolist()
    assert again.diff(oracle.diff(snap, "default", mode=1)) == []


# ------------------------------------------------------------------ second opinion: the Python restatement (explicit Go-slice model)

@pytest.mark.parametrize("seed", range(30))
def test_oracle_agrees_with_python_restatement(garecon, oracle, seed):
    import importlib
    pyref = importlib.import_module("oracle.pyref")
    objects, actual, bindings, known = egbcases.random_bindings(seed) if seed else egbcases.hand_cases()
    snap = garecon.pack(objects, actual)
    want = oracle.bindings_diff(snap, garecon.pack_bindings(bindings, known))
    st, ops = pyref.bindings_diff(objects, actual, bindings, set(known))
    assert st == want.status_ga.tolist()
    assert [tuple(int(x) for x in op) for op in want.ops.tolist()] == ops
