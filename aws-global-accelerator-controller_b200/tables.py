"""Snapshot packer: a plain-Python model of the informer cache + listed AWS resources -> the SoA tables of
include/garecon.h.

This is the small-case twin of the Go-side packer a maintainer would write: it walks
objects exactly the way the reference reads them and lays strings out row-major in one byte slab per
table group, so that consecutive rows are consecutive bytes.

Model (all plain dicts / lists; order is preserved and meaningful):

  objects: [ {kind: "service"|"ingress", ns, name,
              spec_type: "ClusterIP"|"NodePort"|"LoadBalancer"|"ExternalName",   # Service only
              lb_class: bool,                                                    # spec.loadBalancerClass != nil
              ingress_class: None|str,                                           # *spec.ingressClassName
              annotations: {k: v} | [(k, v)...],
              lb_ingress: [hostname...],
              ports: [(number, protocol_str)...] | [number...] } ]
  actual:  {lbs: [{region, name, dns, arn, state}],
            accelerators: [{arn, name, dns, enabled, tags: [(k, v)...],
                            listeners: [{arn, proto: "TCP"|"UDP", ports: [from...],
                                         egs: [{arn, endpoints: [id...]}]}]}],
            zones: [{id, name, records: [{name, type: "A"|"TXT"|"CNAME"|"AAAA"|..., alias: None|str, values: [str...]}]}]}
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

_SPEC = {"ClusterIP": 0, "NodePort": 1, "LoadBalancer": 2, "ExternalName": 3, "": 0}
_LBSTATE = {"active": 0, "provisioning": 1, "active_impaired": 2, "failed": 3}
_RR = {"A": abi.RR_A, "TXT": abi.RR_TXT, "CNAME": abi.RR_CNAME, "AAAA": abi.RR_AAAA}


class _Slab:
    """Byte slab under construction.  put() hands out a handle; finish() lays the bytes out in one of several orders (the
    engine's results must not depend on where strings sit in the slab) and refs() turns handles into gar_str values:
      "row"     insertion order = row-major by parent (an object's key, annotation values, hostname ... are neighbours)
      "level"   column-major: all strings of one column together, columns in first-use order
      "reverse" insertion order reversed
      "shuffle" seeded random order
    sub() is a handle for a slice of an earlier string (obj_ns / obj_name inside the "ns/name" key)."""

    def __init__(self, layout: str = "row", seed: int = 0):
        self.layout, self.seed = layout, seed
        self.items: list[tuple[bytes, str]] = []
        self.subs: list[tuple[int, int, int]] = []
        self.off: list[int] = []
        self.buf = bytearray()

    def put(self, s, col: str = "") -> int:
        b = s if isinstance(s, (bytes, bytearray)) else str(s).encode("utf-8", "surrogatepass")
        self.items.append((bytes(b), col))
        return len(self.items) - 1

    def sub(self, handle: int, start: int, length: int) -> int:
        self.subs.append((handle, start, length))
        return -len(self.subs)

    def finish(self):
        n = len(self.items)
        order = list(range(n))
        if self.layout == "level":
            first: dict[str, int] = {}
            for _, col in self.items:
                first.setdefault(col, len(first))
            order.sort(key=lambda i: first[self.items[i][1]])
        elif self.layout == "reverse":
            order.reverse()
        elif self.layout == "shuffle":
            import random
            random.Random(self.seed).shuffle(order)
        elif self.layout != "row":
            raise ValueError(f"unknown slab layout {self.layout!r}")
        self.off = [0] * n
        for i in order:
            self.off[i] = len(self.buf)
            self.buf += self.items[i][0]

    def ref(self, handle) -> int:
        if handle is None:
            return 0
        if handle < 0:
            h, start, length = self.subs[-handle - 1]
            return (length << abi.OFF_BITS) | (self.off[h] + start)
        return (len(self.items[handle][0]) << abi.OFF_BITS) | self.off[handle]

    def refs(self, handles) -> list[int]:
        return [self.ref(h) for h in handles]


def _ptr(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class Snapshot:
    """Owns the numpy buffers behind a GarObjects / GarActual pair."""

    def __init__(self):
        self.objects = abi.GarObjects()
        self.actual = abi.GarActual()
        self.arrays: dict[str, np.ndarray] = {}

    def _set(self, struct, name, values, dtype, ctype):
        arr = np.ascontiguousarray(np.asarray(values, dtype=dtype))
        if arr.size == 0:
            arr = np.zeros(1, dtype=dtype)[:0].copy()
        self.arrays[name] = arr
        # keep a 1-element backing store for empty arrays so the pointer is never NULL
        if arr.size == 0:
            backing = np.zeros(1, dtype=dtype)
            self.arrays[name + "#backing"] = backing
            setattr(struct, name, _ptr(backing, ctype))
        else:
            setattr(struct, name, _ptr(arr, ctype))

    def obj_str(self, ref: int) -> bytes:
        ref = int(ref)
        off, ln = ref & abi.OFF_MASK, ref >> abi.OFF_BITS
        return bytes(self.arrays["o.slab"][off:off + ln])

    def act_str(self, ref: int) -> bytes:
        ref = int(ref)
        off, ln = ref & abi.OFF_MASK, ref >> abi.OFF_BITS
        return bytes(self.arrays["a.slab"][off:off + ln])

    def input_bytes(self) -> int:
        return sum(a.nbytes for k, a in self.arrays.items() if not k.endswith("#backing"))


def pack(objects: list[dict], actual: dict | None = None, layout: str = "row", seed: int = 0) -> Snapshot:
    """layout: where the strings sit inside the two slabs (see _Slab); the tables are otherwise identical."""
    actual = actual or {}
    snap = Snapshot()
    o = snap.objects
    sl = _Slab(layout, seed)
    kind, spec, flags, ns, name, icls = [], [], [], [], [], []
    ann_b, lbi_b, port_b = [0], [0], [0]
    ann_k, ann_v, lbi_h, port_n, port_p = [], [], [], [], []
    for ob in objects:
        k = abi.KIND_SERVICE if ob.get("kind", "service") == "service" else abi.KIND_INGRESS
        kind.append(k)
        spec.append(_SPEC[ob.get("spec_type", "LoadBalancer" if k == abi.KIND_SERVICE else "")] if k == abi.KIND_SERVICE else 0)
        f = 0
        if ob.get("lb_class"):
            f |= abi.OBJ_HAS_LB_CLASS
        if ob.get("ingress_class") is not None:
            f |= abi.OBJ_HAS_INGRESS_CLASS
        flags.append(f)
        # layout rule: ns and name are slices of one "ns/name" string (the workqueue key, reconcile.go:47)
        nsb = str(ob.get("ns", "default")).encode("utf-8", "surrogatepass")
        nmb = str(ob["name"]).encode("utf-8", "surrogatepass")
        kref = sl.put(nsb + b"/" + nmb, "obj_key")
        ns.append(sl.sub(kref, 0, len(nsb)))
        name.append(sl.sub(kref, len(nsb) + 1, len(nmb)))
        icls.append(sl.put(ob["ingress_class"], "obj_ingress_class") if ob.get("ingress_class") is not None else None)
        anns = ob.get("annotations", {})
        items = list(anns.items()) if isinstance(anns, dict) else list(anns)
        for ak, av in items:
            ann_k.append(sl.put(ak, "ann_key"))
            ann_v.append(sl.put(av, "ann_val"))
        ann_b.append(len(ann_k))
        for h in ob.get("lb_ingress", []):
            lbi_h.append(sl.put(h, "lbi_hostname"))
        lbi_b.append(len(lbi_h))
        for p in ob.get("ports", []):
            if isinstance(p, (tuple, list)):
                port_n.append(int(p[0]))
                port_p.append(sl.put(p[1], "port_proto") if k == abi.KIND_SERVICE else None)
            else:
                port_n.append(int(p))
                port_p.append(sl.put("TCP", "port_proto") if k == abi.KIND_SERVICE else None)
        port_b.append(len(port_n))
    sl.finish()
    ns, name, icls, ann_k, ann_v, lbi_h, port_p = (sl.refs(x) for x in (ns, name, icls, ann_k, ann_v, lbi_h, port_p))
    u8, u32, u64, i32 = np.uint8, np.uint32, np.uint64, np.int32
    o.n_objects = len(objects)
    snap._set(o, "obj_kind", kind, u8, C.c_uint8)
    snap._set(o, "obj_spec_type", spec, u8, C.c_uint8)
    snap._set(o, "obj_flags", flags, u8, C.c_uint8)
    snap._set(o, "obj_ns", ns, u64, C.c_uint64)
    snap._set(o, "obj_name", name, u64, C.c_uint64)
    snap._set(o, "obj_ingress_class", icls, u64, C.c_uint64)
    snap._set(o, "obj_ann_begin", ann_b, u32, C.c_uint32)
    snap._set(o, "obj_lbi_begin", lbi_b, u32, C.c_uint32)
    snap._set(o, "obj_port_begin", port_b, u32, C.c_uint32)
    o.n_ann = len(ann_k)
    snap._set(o, "ann_key", ann_k, u64, C.c_uint64)
    snap._set(o, "ann_val", ann_v, u64, C.c_uint64)
    o.n_lbi = len(lbi_h)
    snap._set(o, "lbi_hostname", lbi_h, u64, C.c_uint64)
    o.n_ports = len(port_n)
    snap._set(o, "port_number", port_n, i32, C.c_int32)
    snap._set(o, "port_proto", port_p, u64, C.c_uint64)
    slab = np.frombuffer(bytes(sl.buf) + b"\0" * 16, dtype=np.uint8).copy()
    snap.arrays["o.slab"] = slab
    o.slab = _ptr(slab, C.c_uint8)
    o.slab_len = len(sl.buf)

    a = snap.actual
    sl = _Slab(layout, seed + 1)
    lbs = actual.get("lbs", [])
    lb_cols = {c: [] for c in ("lb_region", "lb_name", "lb_dns", "lb_arn")}
    for x in lbs:  # row-major: the four strings of one load balancer are neighbours
        for c, f in (("lb_region", "region"), ("lb_name", "name"), ("lb_dns", "dns"), ("lb_arn", "arn")):
            lb_cols[c].append(sl.put(x[f], c))

    accs = actual.get("accelerators", [])
    acc_arn, acc_name, acc_dns, acc_en = [], [], [], []
    tag_b, lis_b, tag_k, tag_v = [0], [0], [], []
    lis_arn, lis_proto, pr_b, eg_b, pr_from = [], [], [0], [0], []
    eg_arn, ep_b, ep_id = [], [0], []
    # strings are laid out accelerator-row-major: accelerator, its tags, its listeners, their egs
    for x in accs:
        acc_arn.append(0)
        acc_name.append(sl.put(x.get("name", ""), "acc_name"))
        acc_dns.append(sl.put(x.get("dns", ""), "acc_dns"))
        acc_en.append(1 if x.get("enabled", True) else 0)
        for tk, tv in x.get("tags", []):
            tag_k.append(sl.put(tk, "tag_key"))
            tag_v.append(sl.put(tv, "tag_val"))
        tag_b.append(len(tag_k))
        for li in x.get("listeners", []):
            lis_arn.append(0)
            lis_proto.append(abi.PROTO_UDP if li.get("proto", "TCP") == "UDP" else abi.PROTO_TCP)
            pr_from.extend(int(p) for p in li.get("ports", []))
            pr_b.append(len(pr_from))
            for eg in li.get("egs", []):
                eg_arn.append(0)
                for e in eg.get("endpoints", []):
                    ep_id.append(sl.put(e, "ep_id"))
                ep_b.append(len(ep_id))
            eg_b.append(len(eg_arn))
        lis_b.append(len(lis_arn))
    zones = actual.get("zones", [])
    z_id, z_name, rec_b = [], [], [0]
    r_name, r_type, r_has, r_alias, val_b, v_val = [], [], [], [], [0], []
    for z in zones:
        z_id.append(0)
        z_name.append(sl.put(z["name"], "zone_name"))
        for r in z.get("records", []):
            r_name.append(sl.put(r["name"], "rec_name"))
            r_type.append(_RR.get(r.get("type", "A"), abi.RR_OTHER))
            alias = r.get("alias")
            r_has.append(1 if alias is not None else 0)
            r_alias.append(sl.put(alias, "rec_alias_dns") if alias is not None else None)
            for v in r.get("values", []):
                v_val.append(sl.put(v, "val_value"))
            val_b.append(len(v_val))
        rec_b.append(len(r_name))
    sl.finish()
    acc_name, acc_dns, tag_k, tag_v, ep_id, z_name, r_name, r_alias, v_val = (
        sl.refs(x) for x in (acc_name, acc_dns, tag_k, tag_v, ep_id, z_name, r_name, r_alias, v_val))
    a.n_lbs = len(lbs)
    for c in lb_cols:
        snap._set(a, c, sl.refs(lb_cols[c]), u64, C.c_uint64)
    snap._set(a, "lb_state", [_LBSTATE[x.get("state", "active")] for x in lbs], u8, C.c_uint8)
    a.n_accels = len(accs)
    snap._set(a, "acc_name", acc_name, u64, C.c_uint64)
    snap._set(a, "acc_dns", acc_dns, u64, C.c_uint64)
    snap._set(a, "acc_enabled", acc_en, u8, C.c_uint8)
    snap._set(a, "acc_tag_begin", tag_b, u32, C.c_uint32)
    snap._set(a, "acc_lis_begin", lis_b, u32, C.c_uint32)
    a.n_tags = len(tag_k)
    snap._set(a, "tag_key", tag_k, u64, C.c_uint64)
    snap._set(a, "tag_val", tag_v, u64, C.c_uint64)
    a.n_listeners = len(lis_arn)
    snap._set(a, "lis_proto", lis_proto, u8, C.c_uint8)
    snap._set(a, "lis_pr_begin", pr_b, u32, C.c_uint32)
    snap._set(a, "lis_eg_begin", eg_b, u32, C.c_uint32)
    a.n_port_ranges = len(pr_from)
    snap._set(a, "pr_from", pr_from, i32, C.c_int32)
    a.n_egs = len(eg_arn)
    snap._set(a, "eg_ep_begin", ep_b, u32, C.c_uint32)
    a.n_endpoints = len(ep_id)
    snap._set(a, "ep_id", ep_id, u64, C.c_uint64)

    a.n_zones = len(zones)
    snap._set(a, "zone_name", z_name, u64, C.c_uint64)
    snap._set(a, "zone_rec_begin", rec_b, u32, C.c_uint32)
    a.n_records = len(r_name)
    snap._set(a, "rec_name", r_name, u64, C.c_uint64)
    snap._set(a, "rec_type", r_type, u8, C.c_uint8)
    snap._set(a, "rec_has_alias", r_has, u8, C.c_uint8)
    snap._set(a, "rec_alias_dns", r_alias, u64, C.c_uint64)
    snap._set(a, "rec_val_begin", val_b, u32, C.c_uint32)
    a.n_values = len(v_val)
    snap._set(a, "val_value", v_val, u64, C.c_uint64)
    slab = np.frombuffer(bytes(sl.buf) + b"\0" * 16, dtype=np.uint8).copy()
    snap.arrays["a.slab"] = slab
    a.slab = _ptr(slab, C.c_uint8)
    a.slab_len = len(sl.buf)
    snap.model = (objects, actual)
    return snap


class Bindings:
    """Owns the buffers behind a GarBindings struct."""

    def __init__(self):
        self.struct = abi.GarBindings()
        self.arrays = {}


def pack_bindings(bindings: list[dict], known_egs: list[str]) -> Bindings:
    """EndpointGroupBinding objects -> gar_bindings.  Each binding:
       {ns, ref: None | ("service"|"ingress", name), eg_arn, deleting: bool, finalizers: bool, observed: bool, endpoint_ids: [arn...]}"""
    out = Bindings()
    b = out.struct
    sl = _Slab()
    flags, kinds, keys, arns, ep_b, eps = [], [], [], [], [0], []
    for x in bindings:
        f = (abi.EGB_DELETING if x.get("deleting") else 0) | (abi.EGB_HAS_FINALIZERS if x.get("finalizers", True) else 0) | (abi.EGB_OBSERVED if x.get("observed", True) else 0)
        flags.append(f)
        ref = x.get("ref")
        kinds.append(0 if ref is None else (1 if ref[0] == "service" else 2))
        keys.append(sl.put(f"{x.get('ns', 'default')}/{ref[1]}") if ref is not None else None)
        arns.append(sl.put(x.get("eg_arn", "")))
        for e in x.get("endpoint_ids", []):
            eps.append(sl.put(e))
        ep_b.append(len(eps))
    known = [sl.put(k) for k in known_egs]
    sl.finish()
    keys, arns, eps, known = sl.refs(keys), sl.refs(arns), sl.refs(eps), sl.refs(known)

    def setcol(name, values, dtype, ctype):
        arr = np.ascontiguousarray(np.asarray(values if len(values) else [0], dtype=dtype))
        out.arrays[name] = arr
        setattr(b, name, _ptr(arr, ctype))

    b.n_bindings = len(bindings)
    setcol("egb_flags", flags, np.uint8, C.c_uint8)
    setcol("egb_ref_kind", kinds, np.uint8, C.c_uint8)
    setcol("egb_ref_key", keys, np.uint64, C.c_uint64)
    setcol("egb_eg_arn", arns, np.uint64, C.c_uint64)
    setcol("egb_ep_begin", ep_b, np.uint32, C.c_uint32)
    b.n_endpoint_ids = len(eps)
    setcol("ep_id", eps, np.uint64, C.c_uint64)
    b.n_known_egs = len(known)
    setcol("known_eg_arn", known, np.uint64, C.c_uint64)
    slab = np.frombuffer(bytes(sl.buf) + b"\0" * 16, dtype=np.uint8).copy()
    out.arrays["slab"] = slab
    b.slab = _ptr(slab, C.c_uint8)
    b.slab_len = len(sl.buf)
    return out


# ------------------------------------------------------------------ numpy-level table tools (sharded mode: building slices
# out of generator chunks, and the union of slices for parity checks)

# table -> (row-count field, [(column, kind)]) with kind: "str" (gar_str, refs into the side's slab), "u8", "i32",
# ("csr", child table) = begin array of n+1 entries into the child table
OBJ_TABLES = {
    "obj": ("n_objects", [("obj_kind", "u8"), ("obj_spec_type", "u8"), ("obj_flags", "u8"), ("obj_ns", "str"), ("obj_name", "str"), ("obj_ingress_class", "str"),
                          ("obj_ann_begin", ("csr", "ann")), ("obj_lbi_begin", ("csr", "lbi")), ("obj_port_begin", ("csr", "port"))]),
    "ann": ("n_ann", [("ann_key", "str"), ("ann_val", "str")]),
    "lbi": ("n_lbi", [("lbi_hostname", "str")]),
    "port": ("n_ports", [("port_number", "i32"), ("port_proto", "str")]),
}
ACT_TABLES = {
    "lb": ("n_lbs", [("lb_region", "str"), ("lb_name", "str"), ("lb_dns", "str"), ("lb_arn", "str"), ("lb_state", "u8")]),
    "acc": ("n_accels", [("acc_name", "str"), ("acc_dns", "str"), ("acc_enabled", "u8"), ("acc_tag_begin", ("csr", "tag")), ("acc_lis_begin", ("csr", "lis"))]),
    "tag": ("n_tags", [("tag_key", "str"), ("tag_val", "str")]),
    "lis": ("n_listeners", [("lis_proto", "u8"), ("lis_pr_begin", ("csr", "pr")), ("lis_eg_begin", ("csr", "eg"))]),
    "pr": ("n_port_ranges", [("pr_from", "i32")]),
    "eg": ("n_egs", [("eg_ep_begin", ("csr", "ep"))]),
    "ep": ("n_endpoints", [("ep_id", "str")]),
    "zone": ("n_zones", [("zone_name", "str"), ("zone_rec_begin", ("csr", "rec"))]),
    "rec": ("n_records", [("rec_name", "str"), ("rec_type", "u8"), ("rec_has_alias", "u8"), ("rec_alias_dns", "str"), ("rec_val_begin", ("csr", "val"))]),
    "val": ("n_values", [("val_value", "str")]),
}
_DT = {"u8": np.uint8, "i32": np.int32, "str": np.uint64}
_CT = {"u8": C.c_uint8, "i32": C.c_int32, "str": C.c_uint64}
ACC_FAMILY = ("acc", "tag", "lis", "pr", "eg", "ep")
REC_FAMILY = ("zone", "rec", "val")


def _view(ptr, n, dtype):
    n = int(n)
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype)


def columns(struct, tables) -> dict:
    """Zero-copy numpy views of every column of a GarObjects / GarActual (plus "slab"); keep the owner alive."""
    cols = {}
    for t, (nf, cl) in tables.items():
        n = getattr(struct, nf)
        for name, kind in cl:
            if isinstance(kind, tuple):
                cols[name] = _view(getattr(struct, name), n + 1, np.uint32)
            else:
                cols[name] = _view(getattr(struct, name), n, _DT[kind])
    cols["slab"] = _view(struct.slab, struct.slab_len, np.uint8)
    return cols


def _rebase(refs: np.ndarray, delta: int) -> np.ndarray:
    return refs + np.uint64(delta)  # the offset lives in the low bits of a gar_str


def from_columns(o_cols: dict, a_cols: dict) -> Snapshot:
    """Snapshot (owning copies) from column dicts as returned by columns()."""
    snap = Snapshot()
    for struct, cols, tables, pre in ((snap.objects, o_cols, OBJ_TABLES, "o."), (snap.actual, a_cols, ACT_TABLES, "a.")):
        for t, (nf, cl) in tables.items():
            n = None
            for name, kind in cl:
                if isinstance(kind, tuple):
                    arr = np.asarray(cols[name], dtype=np.uint32)
                    n = len(arr) - 1
                    snap._set(struct, name, arr, np.uint32, C.c_uint32)
                else:
                    arr = np.asarray(cols[name], dtype=_DT[kind])
                    n = len(arr)
                    snap._set(struct, name, arr, _DT[kind], _CT[kind])
            setattr(struct, nf, int(n))
        slab = np.concatenate([np.asarray(cols["slab"], dtype=np.uint8), np.zeros(64, dtype=np.uint8)])
        snap.arrays[pre + "slab"] = slab
        struct.slab = _ptr(slab, C.c_uint8)
        struct.slab_len = len(slab) - 64
    return snap


def take_families(parts: dict) -> dict:
    """Actual-side columns assembled from different sources: parts = {family tables tuple: a_cols of the source}.  Slabs are
    concatenated and the string references of each family rebased."""
    out, slabs, base = {}, [], 0
    for fam, cols in parts.items():
        for t in fam:
            for name, kind in ACT_TABLES[t][1]:
                out[name] = _rebase(cols[name], base) if kind == "str" else cols[name]
        slabs.append(cols["slab"])
        base += len(cols["slab"])
    out["slab"] = np.concatenate(slabs) if slabs else np.zeros(0, dtype=np.uint8)
    return out


def concat_cols(tables: dict, names, parts: list) -> dict:
    """Column dicts of the tables `names` (one side: OBJ_TABLES or ACT_TABLES) concatenated in list order: rows appended, CSR
    begins shifted, string references rebased into the concatenation of the parts' slabs.  The zone table is the same in every
    part (disjoint record ranges): it is taken once and its begins add up."""
    out = {}
    slab_base = np.concatenate([[0], np.cumsum([len(p["slab"]) for p in parts])]).astype(np.int64)
    for t in names:
        for name, kind in tables[t][1]:
            if isinstance(kind, tuple):
                if t == "zone":
                    out[name] = np.sum([p[name].astype(np.int64) for p in parts], axis=0).astype(np.uint32)
                    continue
                begins, off = [np.zeros(1, dtype=np.int64)], 0
                for p in parts:
                    b = p[name].astype(np.int64)
                    begins.append(b[1:] + off)
                    off += int(b[-1])
                out[name] = np.concatenate(begins).astype(np.uint32)
            elif t == "zone":
                out[name] = _rebase(parts[0][name], 0)
            else:
                arrs = [(_rebase(p[name], int(slab_base[k])) if kind == "str" else p[name]) for k, p in enumerate(parts)]
                out[name] = np.concatenate(arrs) if arrs else np.zeros(0, dtype=_DT[kind])
    out["slab"] = np.concatenate([p["slab"] for p in parts]) if parts else np.zeros(0, dtype=np.uint8)
    return out


def concat_slices(slices) -> Snapshot:
    """The cluster a list of sharded-mode slices [(o_cols, a_cols), ...] (rank order) stands for: every list concatenated in
    rank order, the replicated zone table taken once."""
    return from_columns(concat_cols(OBJ_TABLES, list(OBJ_TABLES), [s[0] for s in slices]), concat_cols(ACT_TABLES, list(ACT_TABLES), [s[1] for s in slices]))


def shard_bases(slices) -> list:
    """GarShard (rank, n_ranks, global row of every table's first row) for each slice [(o_cols, a_cols), ...]."""
    out, g = [], len(slices)
    base = dict(obj=0, lb=0, acc=0, lis=0, eg=0, rec=0, val=0)
    for r, (o, a) in enumerate(slices):
        out.append(abi.GarShard(r, g, base["obj"], base["lb"], base["acc"], base["lis"], base["eg"], base["rec"], base["val"]))
        base["obj"] += len(o["obj_kind"])
        base["lb"] += len(a["lb_state"])
        base["acc"] += len(a["acc_enabled"])
        base["lis"] += len(a["lis_proto"])
        base["eg"] += len(a["eg_ep_begin"]) - 1
        base["rec"] += len(a["rec_type"])
        base["val"] += len(a["val_value"])
    return out
