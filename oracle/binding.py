"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE: importable only from tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs)."""
from __future__ import annotations

import ctypes as C
import importlib
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
abi = importlib.import_module("aws-global-accelerator-controller_b200.abi")

_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "oracle.cpp"
    hdr = HERE.parent / "include" / "garecon.h"
    if force or not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-s", "-C", str(HERE), "-B", "liboracle.so"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        L = C.CDLL(str(LIB))
        L.orc_diff.argtypes = [C.POINTER(abi.GarObjects), C.POINTER(abi.GarActual), C.c_char_p, C.c_int, C.c_int, C.POINTER(C.POINTER(abi.GarChangeset))]
        L.orc_diff.restype = C.c_int
        L.orc_diff_keys.argtypes = [C.POINTER(abi.GarObjects), C.POINTER(abi.GarActual), C.c_char_p, C.c_int, C.POINTER(C.c_uint32), C.c_uint32,
                                    C.POINTER(C.c_uint8), C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.POINTER(abi.GarChangeset))]
        L.orc_diff_keys.restype = C.c_int
        L.orc_bindings_diff.argtypes = [C.POINTER(abi.GarObjects), C.POINTER(abi.GarActual), C.c_char_p, C.POINTER(abi.GarBindings), C.POINTER(C.POINTER(abi.GarChangeset))]
        L.orc_bindings_diff.restype = C.c_int
        L.orc_free.argtypes = [C.POINTER(abi.GarChangeset)]
        L.orc_free.restype = None
        L.orc_detect_cloud_provider.argtypes = [C.c_char_p, C.c_uint32]
        L.orc_get_lb_name_from_hostname.argtypes = [C.c_char_p, C.c_uint32] + [C.POINTER(C.c_uint32)] * 4
        L.orc_parse_listen_ports.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_int32), C.c_int]
        L.orc_json_valid.argtypes = [C.c_char_p, C.c_uint32]
        L.orc_listener_port_changed.argtypes = [C.POINTER(C.c_int32), C.c_uint32, C.POINTER(C.c_int32), C.c_uint32]
        L.orc_service_protocol.argtypes = [C.POINTER(C.c_char_p), C.c_uint32]
        L.orc_parent_domain.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_int]
        L.orc_find_a_record.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_uint8), C.c_uint32, C.c_char_p]
        L.orc_need_records_update.argtypes = [C.c_int, C.c_char_p, C.c_char_p]
        L.orc_route53_owner_value.argtypes = [C.c_char_p] * 4 + [C.c_char_p, C.c_int]
        _lib = L
    return _lib


def diff(snap, cluster: str = "default", mode: int = 1, threads: int = 1):
    """Run the oracle on a packed snapshot; returns abi.ChangeSet (host numpy copies)."""
    L = lib()
    out = C.POINTER(abi.GarChangeset)()
    rc = L.orc_diff(C.byref(snap.objects), C.byref(snap.actual), cluster.encode(), mode, threads, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_diff rc={rc}")
    try:
        return abi.ChangeSet(out.contents)
    finally:
        L.orc_free(out)


def diff_keys(snap, rows, deleted=(), cluster: str = "default", mode: int = 1):
    """Incremental-mode oracle: per-key reconcile of `rows` + processDelete of the deleted keys."""
    L = lib()
    ks = abi.make_keyset(rows, deleted)
    out = C.POINTER(abi.GarChangeset)()
    rc = L.orc_diff_keys(C.byref(snap.objects), C.byref(snap.actual), cluster.encode(), mode, ks.rows, ks.n_rows, ks.deleted_kind, ks.deleted_key, ks.n_deleted, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_diff_keys rc={rc}")
    try:
        return abi.ChangeSet(out.contents, keyset=True)
    finally:
        L.orc_free(out)


def bindings_diff(snap, bindings, cluster: str = "default"):
    L = lib()
    out = C.POINTER(abi.GarChangeset)()
    rc = L.orc_bindings_diff(C.byref(snap.objects), C.byref(snap.actual), cluster.encode(), C.byref(bindings.struct), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_bindings_diff rc={rc}")
    try:
        return abi.ChangeSet(out.contents, keyset=True, bindings=True)
    finally:
        L.orc_free(out)


def diff_raw(objects, actual, cluster: bytes, mode: int, threads: int):
    """Timing entry: runs the oracle and frees the result, returns n_ops."""
    L = lib()
    out = C.POINTER(abi.GarChangeset)()
    rc = L.orc_diff(C.byref(objects), C.byref(actual), cluster, mode, threads, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_diff rc={rc}")
    n = int(out.contents.n_ops)
    L.orc_free(out)
    return n


# ---- unit entry points ------------------------------------------------------------

def detect_cloud_provider(h: str) -> int:
    b = h.encode()
    return lib().orc_detect_cloud_provider(b, len(b))


def get_lb_name_from_hostname(h: str):
    b = h.encode()
    v = [C.c_uint32() for _ in range(4)]
    code = lib().orc_get_lb_name_from_hostname(b, len(b), *[C.byref(x) for x in v])
    if code > 2:
        return code, None, None
    no, nl, ro, rl = [x.value for x in v]
    return code, b[no:no + nl].decode(), b[ro:ro + rl].decode()


def parse_listen_ports(val) -> list[int] | None:
    b = val if isinstance(val, bytes) else val.encode("utf-8", "surrogatepass")
    cap = max(16, len(b))
    out = (C.c_int32 * cap)()
    n = lib().orc_parse_listen_ports(b, len(b), out, cap)
    return None if n < 0 else list(out[:n])


def json_valid(val) -> bool:
    b = val if isinstance(val, bytes) else val.encode("utf-8", "surrogatepass")
    return bool(lib().orc_json_valid(b, len(b)))


def listener_port_changed(lis, des) -> bool:
    a = (C.c_int32 * max(1, len(lis)))(*lis)
    d = (C.c_int32 * max(1, len(des)))(*des)
    return bool(lib().orc_listener_port_changed(a, len(lis), d, len(des)))


def service_protocol(protos: list[str]) -> int:
    arr = (C.c_char_p * max(1, len(protos)))(*[p.encode() for p in protos])
    return lib().orc_service_protocol(arr, len(protos))


def parent_domain(h: str) -> str:
    b = h.encode()
    out = C.create_string_buffer(len(b) + 1)
    n = lib().orc_parent_domain(b, len(b), out, len(b) + 1)
    return out.raw[:n].decode()


def find_a_record(names: list[str], types: list[int], hostname: str) -> int:
    arr = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
    ty = (C.c_uint8 * max(1, len(types)))(*types)
    return lib().orc_find_a_record(arr, ty, len(names), hostname.encode())


def need_records_update(has_alias: bool, alias_dns: str, accel_dns: str) -> bool:
    return bool(lib().orc_need_records_update(int(has_alias), alias_dns.encode(), accel_dns.encode()))


def route53_owner_value(cluster, resource, ns, name) -> str:
    out = C.create_string_buffer(1024)
    n = lib().orc_route53_owner_value(cluster.encode(), resource.encode(), ns.encode(), name.encode(), out, 1024)
    return out.raw[:n].decode()
