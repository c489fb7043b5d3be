"""ctypes binding of libgarecon_synth.so: deterministic synthetic snapshots for tests and bench.py
(workload generation only; see gen.cpp)."""
from __future__ import annotations

import ctypes as C

import numpy as np
from pathlib import Path

from .. import abi

LIB_PATH = Path(__file__).resolve().parent / "libgarecon_synth.so"


class SynthConfig(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_objects", C.c_uint32), ("frac_ingress", C.c_float), ("n_zones", C.c_uint32),
        ("frac_r53", C.c_float), ("min_hostnames", C.c_uint32), ("max_hostnames", C.c_uint32), ("frac_wildcard", C.c_float),
        ("svc_ports", C.c_uint32), ("frac_hot", C.c_float), ("hot_pool", C.c_uint32), ("frac_listen_ann", C.c_float),
        ("frac_unmanaged", C.c_float), ("frac_ineligible", C.c_float),
        ("p_missing_acc", C.c_float), ("p_port_drift", C.c_float), ("p_proto_drift", C.c_float), ("p_tag_drift", C.c_float),
        ("p_missing_listener", C.c_float), ("p_missing_eg", C.c_float), ("p_lb_not_active", C.c_float), ("p_orphan_acc", C.c_float),
        ("p_rec_missing", C.c_float), ("p_alias_drift", C.c_float), ("p_orphan_rec", C.c_float), ("p_dup_ports", C.c_float),
        ("intern_keys", C.c_uint32), ("index_base", C.c_uint32), ("zone_base", C.c_uint32), ("zones_total", C.c_uint32), ("layout", C.c_uint32), ("emit_mask", C.c_uint32),
        ("cluster", C.c_char * 64),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} not built - run __graft_entry__.build()")
        L = C.CDLL(str(LIB_PATH))
        L.gsyn_preset.argtypes = [C.c_int, C.c_uint32, C.POINTER(SynthConfig)]
        L.gsyn_preset.restype = None
        L.gsyn_generate.argtypes = [C.POINTER(SynthConfig)]
        L.gsyn_generate.restype = C.c_void_p
        L.gsyn_objects.argtypes = [C.c_void_p]
        L.gsyn_objects.restype = C.POINTER(abi.GarObjects)
        L.gsyn_actual.argtypes = [C.c_void_p]
        L.gsyn_actual.restype = C.POINTER(abi.GarActual)
        L.gsyn_free.argtypes = [C.c_void_p]
        L.gsyn_free.restype = None
        _lib = L
    return _lib


def preset(cfg: int, n: int, **overrides) -> SynthConfig:
    c = SynthConfig()
    lib().gsyn_preset(cfg, n, C.byref(c))
    for k, v in overrides.items():
        setattr(c, k, v.encode() if k == "cluster" else v)
    return c


class SynthSnapshot:
    """Generator-owned snapshot; `.objects` / `.actual` are the C structs (host pointers)."""

    def __init__(self, cfg: SynthConfig):
        self.cfg = cfg
        self._h = lib().gsyn_generate(C.byref(cfg))
        if not self._h:
            raise MemoryError("gsyn_generate failed")
        self.objects = lib().gsyn_objects(self._h).contents
        self.actual = lib().gsyn_actual(self._h).contents

    @property
    def cluster(self) -> str:
        return self.cfg.cluster.decode()

    def close(self):
        if self._h:
            lib().gsyn_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate(cfg: int, n: int, **overrides) -> SynthSnapshot:
    return SynthSnapshot(preset(cfg, n, **overrides))


def cluster_slices(cfg_id: int, n_total: int, n_ranks: int, ranks=None, seed=None, layout: int = 0, n_chunks: int | None = None, threads: int = 1, **overrides):
    """Slices of ONE synthetic cluster of n_total objects for sharded mode (include/garecon.h "sharded mode").

    The cluster is generated as n_chunks chunks (default: one per rank; same seed, cluster-wide object indexes, disjoint zone
    ranges of one zone table), n_chunks a multiple of n_ranks.  Slice r holds the objects and record sets of its own group of
    n_chunks / n_ranks consecutive chunks, but the accelerators of the next rank's group and the load balancers of the group
    after that: nothing an object needs is on its own rank, as after an arbitrary cut of the lists.  Returns [(o_cols, a_cols)]
    for `ranks` (default: all) as numpy column dicts (tables.columns layout); row counts of every slice are a pure function of
    (cfg, n_total, n_ranks, n_chunks), so ranks can generate only their own slice.  `threads` > 1 generates chunks
    concurrently (the generator runs outside the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    from .. import tables
    n_chunks = n_ranks if n_chunks is None else int(n_chunks)
    if n_chunks % n_ranks or n_total % n_chunks:
        raise ValueError("n_chunks must be a multiple of n_ranks and divide n_total")
    cpr = n_chunks // n_ranks
    n_chunk = n_total // n_chunks
    want = list(range(n_ranks)) if ranks is None else list(ranks)

    def cols(job):
        c, mask = job
        cfg = preset(cfg_id, n_chunk, **overrides)
        if seed is not None:
            cfg.seed = seed
        cfg.index_base = c * n_chunk
        cfg.zone_base = c * cfg.n_zones
        cfg.zones_total = n_chunks * cfg.n_zones
        cfg.emit_mask = mask
        cfg.layout = layout  # 0 = strings row-major by parent, 1 = column-major slabs
        snap = SynthSnapshot(cfg)
        return snap, tables.columns(snap.objects, tables.OBJ_TABLES), tables.columns(snap.actual, tables.ACT_TABLES)

    def group(r, shift):
        return [((r + shift) % n_ranks) * cpr + k for k in range(cpr)]

    jobs = []
    for r in want:
        jobs += [(c, 1 | 8) for c in group(r, 0)] + [(c, 4) for c in group(r, 1)] + [(c, 2) for c in group(r, 2)]
    jobs = list(dict.fromkeys(jobs))
    if threads > 1 and len(jobs) > 1:
        with ThreadPoolExecutor(max_workers=min(threads, len(jobs))) as pool:
            done = dict(zip(jobs, pool.map(cols, jobs)))
    else:
        done = {j: cols(j) for j in jobs}

    out = []
    for r in want:
        own = [done[(c, 1 | 8)] for c in group(r, 0)]
        acc_src = [done[(c, 4)] for c in group(r, 1)]
        lb_src = [done[(c, 2)] for c in group(r, 2)]
        o = tables.concat_cols(tables.OBJ_TABLES, list(tables.OBJ_TABLES), [x[1] for x in own])
        a = tables.take_families({tables.REC_FAMILY: tables.concat_cols(tables.ACT_TABLES, tables.REC_FAMILY, [x[2] for x in own]),
                                  tables.ACC_FAMILY: tables.concat_cols(tables.ACT_TABLES, tables.ACC_FAMILY, [x[2] for x in acc_src]),
                                  ("lb",): tables.concat_cols(tables.ACT_TABLES, ("lb",), [x[2] for x in lb_src])})
        out.append(({k: np.array(v) for k, v in o.items()}, {k: np.array(v) for k, v in a.items()}))
    return out
