"""ctypes binding of libgarecon_synth.so: deterministic synthetic snapshots for tests and bench.py
(workload generation only; see gen.cpp)."""
from __future__ import annotations

import ctypes as C
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
        ("intern_keys", C.c_uint32),
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
