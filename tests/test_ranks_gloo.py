"""CPU tier: the N>1 host logic of bench.py (independent clusters per rank, barrier, max over ranks, gathered counters)
with world_size 2 on the gloo backend.  No GPU: each rank diffs its own small snapshot with the oracle."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent('''
    import importlib, json, sys, time
    sys.path.insert(0, %r)
    import __graft_entry__ as ge
    ranks = importlib.import_module("aws-global-accelerator-controller_b200.ranks")
    synth = importlib.import_module("aws-global-accelerator-controller_b200.synth")
    ob = importlib.import_module("oracle.binding")
    R = ranks.Ranks(backend="gloo")
    cfg = synth.preset(2, 2000)
    cfg.seed = ranks.rank_seed(cfg.seed, R.rank)
    snap = synth.SynthSnapshot(cfg)
    R.barrier()
    t0 = time.perf_counter()
    cs = ob.diff(snap, "default", mode=1)
    dt = time.perf_counter() - t0 + 0.01 * R.rank          # rank 1 is artificially slower
    mx = R.max_over_ranks(dt)
    counts = R.gather_counts([snap.objects.n_objects, len(cs.ops), cs.checksum() %% (1 << 62)])
    if R.rank == 0:
        print(json.dumps({"world": R.world, "max": mx, "mine": dt, "counts": counts}))
    R.close()
''') % str(REPO)


def test_two_ranks_gloo(tmp_path):
    import __graft_entry__ as ge
    ge.build_synth()
    ge.build_oracle()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2
    assert d["max"] >= d["mine"] and d["max"] >= 0.01      # the slower rank defines the step time
    (n0, ops0, ck0), (n1, ops1, ck1) = d["counts"]
    assert n0 == n1 == 2000                                  # weak scaling: same work per rank
    assert ck0 != ck1 and ops0 > 0 and ops1 > 0             # but different clusters (rank-specific seed)


def test_rank_seed_is_rank_specific():
    import importlib
    ranks = importlib.import_module("aws-global-accelerator-controller_b200.ranks")
    assert len({ranks.rank_seed(7, r) for r in range(8)}) == 8
    assert ranks.rank_env() == (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))
