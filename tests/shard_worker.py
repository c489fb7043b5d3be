"""Worker of the multi-process sharded-mode tests (launched under torch.distributed.run).

  python shard_worker.py gloo  randmodel <seed> <n_objects>     hostsim engines, CPU tensors, oracle as the arbiter
  python shard_worker.py nccl  synth <cfg> <n_total>            real engines, one GPU per rank, blobs over NCCL; the arbiter is
                                                                 the unsharded GPU diff of the union on rank 0
Prints one JSON line on rank 0."""
import importlib
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

garecon = importlib.import_module("aws-global-accelerator-controller_b200")
shard = importlib.import_module("aws-global-accelerator-controller_b200.shard")
synth = importlib.import_module("aws-global-accelerator-controller_b200.synth")


class Part:
    pass


def main():
    backend, source = sys.argv[1], sys.argv[2]
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        dist.init_process_group("nccl", device_id=dev)
        lib, device = None, local_rank
    else:
        dev = "cpu"
        dist.init_process_group("gloo")
        lib, device = garecon.abi.load_library(ge.build_hostsim()), 0
    rank, world = dist.get_rank(), dist.get_world_size()
    if source == "randmodel":
        import randmodel
        objects, actual = randmodel.make(int(sys.argv[3]), n_objects=int(sys.argv[4]))
        objs_r, act_r, sh = shard.slice_model(objects, actual, world)[rank]
        if len(sys.argv) > 5 and sys.argv[5] == "badzone" and rank == world - 1:  # contract violation on ONE rank only
            act_r = dict(act_r, zones=[dict(act_r["zones"][0], name="other.example.org.")] + act_r["zones"][1:])
        snap = garecon.pack(objs_r, act_r)
        n_total = len(objects)
    else:
        n_total = int(sys.argv[4]) - int(sys.argv[4]) % world
        slices = synth.cluster_slices(int(sys.argv[3]), n_total, world)  # every rank builds all slices (small sizes only)
        sh = garecon.tables.shard_bases(slices)[rank]
        snap = garecon.tables.from_columns(*slices[rank])
    e = garecon.Engine(cluster_name="default", lib=lib, device=device)
    e.load(snap)
    x = (shard.PeerExchange if len(sys.argv) > 5 and sys.argv[5] == "peers" else shard.DistExchange)(e, sh, dev)
    if len(sys.argv) > 5 and sys.argv[5] == "badzone":
        # every rank must raise (the failing one its own error, the others the agreed abort): nobody may hang in a collective
        try:
            x.run()
            raised = ""
        except Exception as ex:  # noqa: BLE001
            raised = type(ex).__name__ + ": " + str(ex)[:80]
        allr = [None] * world
        dist.all_gather_object(allr, raised)
        if rank == 0:
            print(json.dumps({"raised": allr}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    x.run()
    part = e.diff()
    payload = dict(obj_gid=part.obj_gid, status_ga=part.status_ga, status_r53=part.status_r53, derived=part.derived, ops=part.ops,
                   section_begin=part.section_begin, sent=x.bytes_sent, launches=part.kernel_launches)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank == 0:
        parts = []
        for g in gathered:
            p = Part()
            p.__dict__.update(g)
            parts.append(p)
        got = shard.merge_changesets(parts, n_total)
        if source == "randmodel":
            ob = importlib.import_module("oracle.binding")
            want = ob.diff(garecon.pack(objects, actual), "default", mode=1)
        else:
            union = garecon.tables.concat_slices(slices)
            with garecon.Engine(cluster_name="default", device=device) as u:
                u.load(union)
                want = u.diff()
        ok = (np.array_equal(got["status_ga"], want.status_ga) and np.array_equal(got["status_r53"], want.status_r53)
              and np.array_equal(got["derived"], want.derived) and np.array_equal(got["ops"], want.ops)
              and got["section_begin"].tolist() == want.section_begin.tolist())
        print(json.dumps({"ok": bool(ok), "world": world, "n_objects": n_total, "n_ops": int(len(want.ops)), "homed": [int(len(p.obj_gid)) for p in parts],
                          "sent": [int(p.sent) for p in parts], "launches": [int(p.launches) for p in parts]}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
