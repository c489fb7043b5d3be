"""pyref.py — second, independent restatement of the reference's decision logic (pure Python, dict model).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).  It exists because most of the path has no reference
test to pin the C++ oracle against (SURVEY.md §8c "parity unpinned"): two restatements written separately,
in different languages and styles (this one uses `re` for the hostname grammar and Python's json tokenizer
for the listen-ports annotation), are diffed on randomized small snapshots in tests/test_oracle_crosscheck.py.

It evaluates the *dict model* of tables.pack() (not the packed tables), per key, with the reference's own
linear scans.  Citations are reference file:line.
"""
from __future__ import annotations

import json
import re

NONE = 0xFFFFFFFF
PENDING = 0xFFFFFFFE  # GAR_PENDING: a resource an earlier op of the same object creates (include/garecon.h)

ANN_MANAGED = "aws-global-accelerator-controller.h3poteto.dev/global-accelerator-managed"
ANN_R53 = "aws-global-accelerator-controller.h3poteto.dev/route53-hostname"
ANN_IPPRESERVE = "aws-global-accelerator-controller.h3poteto.dev/client-ip-preservation"
ANN_NAME = "aws-global-accelerator-controller.h3poteto.dev/global-accelerator-name"
ANN_TAGS = "aws-global-accelerator-controller.h3poteto.dev/global-accelerator-tags"
ANN_IPTYPE = "aws-global-accelerator-controller.h3poteto.dev/ip-address-type"
ANN_LBTYPE = "service.beta.kubernetes.io/aws-load-balancer-type"
ANN_INGRESS_CLASS = "kubernetes.io/ingress.class"
ANN_LISTEN_PORTS = "alb.ingress.kubernetes.io/listen-ports"
TAG_MANAGED = "aws-global-accelerator-controller-managed"
TAG_OWNER = "aws-global-accelerator-owner"
TAG_HOST = "aws-global-accelerator-target-hostname"
TAG_CLUSTER = "aws-global-accelerator-cluster"

ST_IGNORED, ST_OK, ST_SKIP_NO_LB, ST_REQUEUE_30S, ST_REQUEUE_60S, ST_ERR_RETRY, ST_ERR_NORETRY, ST_PANIC = range(8)
EV_CREATED, EV_DELETED = 1, 2


def status(st, detail=0, ev=0):
    return st | (detail << 8) | (ev << 16)


def head(op, ctrl, kind):
    return op | (ctrl << 8) | (kind << 16)


# ---------------------------------------------------------------- provider.go:8-17, load_balancer.go:32-93

_ALB = re.compile(r"\.elb\.amazonaws\.com\Z")
_NLB = re.compile(r"\.elb\..+\.amazonaws\.com\Z")  # '.' excludes \n, as in Go
_INTERNAL = re.compile(r"\Ainternal-")
_INTERNAL_NAME = re.compile(r"\Ainternal\-([\w\-]+)\-[\w]+\Z", re.ASCII)
_NAME = re.compile(r"\A([\w\-]+)\-[\w]+\Z", re.ASCII)


def tokenise(hostname: str):
    """-> (code, name, region); codes as gar_tok_code."""
    parts = hostname.split(".")
    if len(parts) < 2:
        return (4, None, None)  # panic
    if parts[-2] + "." + parts[-1] != "amazonaws.com":
        return (3, None, None)
    if _ALB.search(hostname):
        sub, region = parts[0], parts[1]
        if _INTERNAL.search(sub):
            m = _INTERNAL_NAME.findall(sub)
            if len(m) != 1:
                return (6, None, None)
            return (0, m[0], region)
        m = _NAME.findall(sub)
        if len(m) != 1:
            return (7, None, None)
        return (1, m[0], region)
    if _NLB.search(hostname):
        sub, region = parts[0], parts[2]
        m = _NAME.findall(sub)
        if len(m) != 1:
            return (8, None, None)
        return (2, m[0], region)
    return (5, None, None)


# ---------------------------------------------------------------- listenerForIngress annotation branch (:526-542)

class _Float:
    pass


class _TypeErr(Exception):
    pass


def _fold(s: str) -> str:
    out = []
    for ch in s:
        if "a" <= ch <= "z":
            out.append(ch.upper())
        elif ch == "ſ":
            out.append("S")
        elif ch == "K":
            out.append("K")
        else:
            out.append(ch)
    return "".join(out)


def _bad_const(_):
    raise ValueError("NaN/Infinity are not JSON")


def parse_listen_ports(val: str):
    """-> list of int32 ports, or None on any json error (caller then uses [])."""
    try:
        doc = json.loads(val, parse_float=lambda s: _Float(), parse_constant=_bad_const, object_pairs_hook=lambda p: ("obj", p))
    except (ValueError, RecursionError):
        return None
    if doc is None:
        return []
    if not isinstance(doc, list):
        return None
    ports = []
    try:
        for el in doc:
            http = https = 0
            if el is None:
                pass
            elif isinstance(el, tuple) and el[0] == "obj":
                for k, v in el[1]:
                    f = _fold(k)
                    if f not in ("HTTP", "HTTPS"):
                        continue
                    if v is None:
                        continue
                    if isinstance(v, bool) or not isinstance(v, int):
                        raise _TypeErr()
                    if not (-(1 << 63) <= v < (1 << 63)):
                        raise _TypeErr()
                    if f == "HTTP":
                        http = v
                    else:
                        https = v
            else:
                raise _TypeErr()
            for x in (http, https):
                if x != 0:
                    x &= 0xFFFFFFFF
                    ports.append(x - (1 << 32) if x >= (1 << 31) else x)
    except _TypeErr:
        return None
    return ports


# ---------------------------------------------------------------- helpers

def _ann(ob):
    a = ob.get("annotations", {})
    return dict(a) if not isinstance(a, dict) else a


def _resource(ob):
    return "service" if ob.get("kind", "service") == "service" else "ingress"


def was_lb_service(ob):  # globalaccelerator/service.go:18-26
    if ob.get("spec_type", "LoadBalancer") == "LoadBalancer":
        return ANN_LBTYPE in _ann(ob) or bool(ob.get("lb_class"))
    return False


def was_alb_ingress(ob):  # ingress.go:19-27
    return ob.get("ingress_class") == "alb" or ANN_INGRESS_CLASS in _ann(ob)


def _ports(ob):
    out = []
    for p in ob.get("ports", []):
        out.append((int(p[0]), p[1]) if isinstance(p, (tuple, list)) else (int(p), "TCP"))
    return out


def desired_listener(ob):
    """-> (ports, proto(0 tcp/1 udp), from_annotation)"""
    if _resource(ob) == "service":
        proto = 0
        for _, pr in _ports(ob):
            low = pr.lower()
            if low == "udp":
                proto = 1
            elif low == "tcp":
                proto = 0
        return [n for n, _ in _ports(ob)], proto, False
    a = _ann(ob)
    if ANN_LISTEN_PORTS in a:
        p = parse_listen_ports(a[ANN_LISTEN_PORTS])
        return (p or []), 0, True
    return [n for n, _ in _ports(ob)], 0, False


def port_changed(lis_ports, des_ports):  # :458-492
    count = {}
    for p in list(lis_ports) + list(des_ports):
        count[p] = count.get(p, 0) + 1
    return any(v <= 1 for v in count.values())


def accelerator_tags(ob):  # :35-51
    out = []
    for piece in _ann(ob).get(ANN_TAGS, "").split(","):
        t = piece.split("=")
        if len(t) == 2:
            out.append((t[0], t[1]))
    return out


def accelerator_name(ob):  # :53-60
    n = _ann(ob).get(ANN_NAME, "")
    return n if n != "" else f"{_resource(ob)}-{ob.get('ns', 'default')}-{ob['name']}"


def tags_contain(actual_tags, target: dict):  # :559-570
    actual = {}
    for k, v in actual_tags:
        actual[k] = v
    return all(actual.get(k, "") == v for k, v in target.items())


def parent_domain(h):  # route53.go:383-386
    return ".".join(h.split(".")[1:])


def owner_value(cluster, resource, ns, name):  # route53.go:18-20
    return f'"heritage=aws-global-accelerator-controller,cluster={cluster},{resource}/{ns}/{name}"'


# ---------------------------------------------------------------- flat row numbering (matches tables.pack order)

class _Rows:
    def __init__(self, actual):
        self.acc, self.lis, self.eg, self.zone, self.rec, self.val = [], [], [], [], [], []
        for ai, acc in enumerate(actual.get("accelerators", [])):
            self.acc.append(acc)
            acc["_row"] = ai
            acc["_lis"] = []
            for li in acc.get("listeners", []):
                li["_row"] = len(self.lis)
                self.lis.append(li)
                acc["_lis"].append(li)
                li["_egs"] = []
                for eg in li.get("egs", []):
                    eg["_row"] = len(self.eg)
                    self.eg.append(eg)
                    li["_egs"].append(eg)
        for zi, z in enumerate(actual.get("zones", [])):
            z["_row"] = zi
            z["_recs"] = []
            for r in z.get("records", []):
                r["_row"] = len(self.rec)
                self.rec.append(r)
                z["_recs"].append(r)
                r["_vals"] = []
                for v in r.get("values", []):
                    r["_vals"].append((len(self.val), v))
                    self.val.append((r, v))
        for i, lb in enumerate(actual.get("lbs", [])):
            lb["_row"] = i


def diff(objects, actual, cluster):
    """-> dict(status_ga, status_r53, derived, ops(list of 6-tuples), section_begin, tok(list of (code,name,region)), dports(list per object|None))"""
    actual = actual or {}
    _Rows(actual)
    accs = actual.get("accelerators", [])
    lbs = actual.get("lbs", [])
    zones = actual.get("zones", [])

    def list_by_resource(resource, ns, name):  # :87-110
        t = {TAG_MANAGED: "true", TAG_OWNER: f"{resource}/{ns}/{name}", TAG_CLUSTER: cluster}
        return [a for a in accs if tags_contain(a.get("tags", []), t)]

    def list_by_hostname(h):  # :62-85
        t = {TAG_MANAGED: "true", TAG_HOST: h, TAG_CLUSTER: cluster}
        return [a for a in accs if tags_contain(a.get("tags", []), t)]

    def delete_chain(ops, objrow, kind, acc):  # :254-288
        lis = eg = NONE
        if len(acc["_lis"]) == 1:
            lis = acc["_lis"][0]["_row"]
            if len(acc["_lis"][0]["_egs"]) == 1:
                eg = acc["_lis"][0]["_egs"][0]["_row"]
        ops.append((head(7, 0, kind if objrow != NONE else 0), objrow, 0, acc["_row"], lis, eg))

    def owned_alias_sets(z, ov):  # route53.go:216-238
        names = []
        for r in z["_recs"]:
            for vi, v in r["_vals"]:
                if v == ov:
                    names.append((r["name"], vi))
        out = []
        for r in z["_recs"]:
            if r.get("alias") is None:
                continue
            for n, vi in names:
                if n == r["name"]:
                    out.append((r, vi))
                    break
        return out

    def owned_metadata_sets(z, ov):  # route53.go:167-181
        return [(r, vi) for r in z["_recs"] for vi, v in r["_vals"] if v == ov]

    def cleanup_records(ops, objrow, kind, ov):  # route53.go:132-165
        h = head(10, 1, kind if objrow != NONE else 0)
        for z in zones:
            for r, vi in owned_alias_sets(z, ov):
                ops.append((h, objrow, 0, z["_row"], r["_row"], vi))
            for r, vi in owned_metadata_sets(z, ov):
                ops.append((h, objrow, 1, z["_row"], r["_row"], vi))

    def hosted_zone(hostname):  # route53.go:335-358
        t = hostname
        while True:
            if t == "":
                return None
            for z in zones:
                if z["name"] == t + ".":
                    return z
            t = parent_domain(t)

    st_ga, st_r53, derived, ga_ops, r53_ops, tok, dports = [], [], [], [], [], [], []
    for row, ob in enumerate(objects):
        kind = 0 if _resource(ob) == "service" else 1
        resource = _resource(ob)
        ann = _ann(ob)
        ns, name = ob.get("ns", "default"), ob["name"]
        hosts = ob.get("lb_ingress", [])
        for h in hosts:
            tok.append(tokenise(h))
        ports, proto, from_ann = desired_listener(ob)
        ga_el = was_lb_service(ob) if kind == 0 else was_alb_ingress(ob)
        r53_el = was_lb_service(ob) if kind == 0 else True
        dv = (1 if proto == 1 else 0) | (2 if ann.get(ANN_IPPRESERVE, "") == "true" else 0) | (4 if ann.get(ANN_IPTYPE, "") in ("ipv4", "IPV4") else 0)
        dv |= (8 if from_ann else 0) | (16 if ga_el else 0) | (32 if ANN_MANAGED in ann else 0) | (64 if r53_el else 0) | (128 if ANN_R53 in ann else 0)
        derived.append(dv)
        dports.append(ports if from_ann else None)

        # ---- globalaccelerator controller (service.go:54-126, ingress.go:56-130)
        def ga():
            if not ga_el:
                return status(ST_IGNORED)
            if len(hosts) < 1:
                return status(ST_SKIP_NO_LB)
            if ANN_MANAGED not in ann:
                for acc in list_by_resource(resource, ns, name):
                    delete_chain(ga_ops, row, kind, acc)
                return status(ST_OK, 0, EV_DELETED)
            # The reference mutates AWS between the lbIngress iterations and re-lists (ListGlobalAcceleratorByResource sees
            # the accelerator it created one iteration earlier; updateEndpointGroup REPLACES the endpoint list).  A batch
            # evaluates a frozen snapshot, so the object carries an overlay of what its own earlier ops changed: deep copies
            # of the accelerators it touched plus the accelerator it created (rows = PENDING).
            import copy
            ev = 0
            over = {}     # accelerator row -> object-local copy
            created = []  # accelerators created by this object's own ops

            def view():
                return [over.get(a["_row"], a) for a in accs] + created

            def own(acc):
                if acc["_row"] == PENDING or acc["_row"] in over:
                    return acc
                c = copy.deepcopy(acc)
                c["_lis"] = [dict(li, _egs=[dict(eg) for eg in li["_egs"]]) for li in acc["_lis"]]
                over[acc["_row"]] = c
                return c

            def desired_tags(lbdns, with_cluster):  # createAccelerator (:654-675) / updateAccelerator (:720-735)
                t = [(TAG_MANAGED, "true"), (TAG_OWNER, f"{resource}/{ns}/{name}"), (TAG_HOST, lbdns)]
                if with_cluster:
                    t.append((TAG_CLUSTER, cluster))
                return t + accelerator_tags(ob)

            for j, h in enumerate(hosts):
                code, lbname, region = tokenise(h)
                if code == 4:
                    return status(ST_PANIC, 0, ev)
                if code == 3:
                    continue
                if code >= 5:
                    return status(ST_ERR_RETRY, 1 + (code - 5), ev)
                lb = next((x for x in lbs if x["region"] == region and x["name"] == lbname), None)  # load_balancer.go:13-30
                if lb is None:
                    return status(ST_ERR_RETRY, 5, ev)
                if lb["dns"] != h:
                    return status(ST_ERR_RETRY, 6, ev)
                if lb.get("state", "active") != "active":
                    return status(ST_REQUEUE_30S, 0, ev)
                t = {TAG_MANAGED: "true", TAG_OWNER: f"{resource}/{ns}/{name}", TAG_CLUSTER: cluster}
                found = [a for a in view() if tags_contain(a.get("tags", []), t)]  # ListGlobalAcceleratorByResource on the overlaid state
                if not found:
                    ga_ops.append((head(1, 0, kind), row, j, lb["_row"], NONE, NONE))
                    ev |= EV_CREATED
                    created.append({"_row": PENDING, "name": accelerator_name(ob), "enabled": True, "tags": desired_tags(lb["dns"], True),
                                    "_lis": [{"_row": PENDING, "proto": "UDP" if proto == 1 else "TCP", "ports": list(ports),
                                              "_egs": [{"_row": PENDING, "endpoints": [lb["arn"]]}]}]})
                    continue
                for acc in found:  # updateGlobalAcceleratorFor* (:290-410)
                    changed = (not acc.get("enabled", True)) or acc.get("name", "") != accelerator_name(ob)
                    if not changed:
                        target = {TAG_MANAGED: "true", TAG_OWNER: f"{resource}/{ns}/{name}", TAG_HOST: lb["dns"]}
                        for k, v in accelerator_tags(ob):
                            target[k] = v
                        changed = not tags_contain(acc.get("tags", []), target)
                    if changed:
                        ga_ops.append((head(2, 0, kind), row, j, acc["_row"], lb["_row"], NONE))
                        acc = own(acc)
                        acc["enabled"], acc["name"] = True, accelerator_name(ob)
                        acc["tags"] = list(acc.get("tags", [])) + desired_tags(lb["dns"], False)  # TagResource: later entries win
                    ls = acc["_lis"]
                    if len(ls) > 1:
                        return status(ST_ERR_RETRY, 7, ev)
                    if len(ls) == 0:
                        ga_ops.append((head(3, 0, kind), row, j, acc["_row"], NONE, NONE))
                        ga_ops.append((head(5, 0, kind), row, j, acc["_row"], NONE, lb["_row"]))
                        acc = own(acc)
                        acc["_lis"] = [{"_row": PENDING, "proto": "UDP" if proto == 1 else "TCP", "ports": list(ports),
                                        "_egs": [{"_row": PENDING, "endpoints": [lb["arn"]]}]}]
                        continue
                    li = ls[0]
                    lproto = 1 if li.get("proto", "TCP") == "UDP" else 0
                    proto_changed = (lproto != proto) if kind == 0 else (lproto != 0)
                    if proto_changed or port_changed(li.get("ports", []), ports):
                        ga_ops.append((head(4, 0, kind), row, j, acc["_row"], li["_row"], NONE))
                        acc = own(acc)
                        li = acc["_lis"][0]
                        li["proto"], li["ports"] = ("UDP" if proto == 1 else "TCP") if kind == 0 else "TCP", list(ports)
                    egs = li["_egs"]
                    if len(egs) > 1:
                        return status(ST_ERR_RETRY, 8, ev)
                    if len(egs) == 0:
                        ga_ops.append((head(5, 0, kind), row, j, acc["_row"], li["_row"], lb["_row"]))
                        acc = own(acc)
                        acc["_lis"][0]["_egs"] = [{"_row": PENDING, "endpoints": [lb["arn"]]}]
                        continue
                    if lb["arn"] not in egs[0].get("endpoints", []):
                        ga_ops.append((head(6, 0, kind), row, j, acc["_row"], egs[0]["_row"], lb["_row"]))
                        acc = own(acc)
                        acc["_lis"][0]["_egs"][0]["endpoints"] = [lb["arn"]]  # updateEndpointGroup replaces the list (:987-1002)
            return status(ST_OK, 0, ev)

        st_ga.append(ga())

        # ---- route53 controller (route53/service.go:48-111, ingress.go:40-104)
        def r53():
            if not r53_el:
                return status(ST_IGNORED)
            ov = owner_value(cluster, resource, ns, name)
            if ANN_R53 not in ann:
                cleanup_records(r53_ops, row, kind, ov)
                return status(ST_OK, 0, EV_DELETED)
            hostnames = ann[ANN_R53].split(",")
            ev = 0
            # object-local overlay of Route53: record sets this object's own earlier ops created (rows = PENDING) or re-pointed
            new_recs = {}   # zone row -> [record dicts]
            new_alias = {}  # record row -> alias DNS name written by an UPSERT of this object

            def alias_sets(z):
                names = []
                allrecs = z["_recs"] + new_recs.get(z["_row"], [])
                for r in allrecs:
                    for vi, v in r["_vals"]:
                        if v == ov:
                            names.append((r["name"], vi))
                out = []
                for r in allrecs:
                    if r.get("alias") is None:
                        continue
                    for n, vi in names:
                        if n == r["name"]:
                            out.append((r, vi))
                            break
                return out

            for j, h in enumerate(hosts):
                code, _, _ = tokenise(h)
                if code == 4:
                    return status(ST_PANIC, 0, ev)
                if code == 3:
                    continue
                if code >= 5:
                    return status(ST_ERR_RETRY, 1 + (code - 5), ev)
                found = list_by_hostname(h)
                if len(found) > 1:
                    return status(ST_REQUEUE_60S, 10, ev)
                if len(found) == 0:
                    return status(ST_REQUEUE_60S, 11, ev)
                acc = found[0]
                created = False
                for k, hn in enumerate(hostnames):
                    z = hosted_zone(hn)
                    if z is None:
                        return status(ST_ERR_RETRY, 9, ev)
                    rec = None
                    for r, _vi in alias_sets(z):  # findARecord (:360-367) over FindOwneredARecordSets on the overlaid state
                        if r.get("type", "A") == "A" and r["name"].replace("\\052", "*", 1) == hn + ".":
                            rec = r
                            break
                    want = acc.get("dns", "") + "."
                    if rec is None:
                        r53_ops.append((head(8, 1, kind), row, (j << 20) | k, z["_row"], acc["_row"], NONE))
                        created = True
                        rn = (hn + ".").replace("*", "\\052")  # Route53 stores '*' escaped
                        new_recs.setdefault(z["_row"], []).extend([
                            {"_row": PENDING, "name": rn, "type": "TXT", "_vals": [(PENDING, ov)]},
                            {"_row": PENDING, "name": rn, "type": "A", "alias": want, "_vals": []}])
                    else:
                        cur = new_alias.get(rec["_row"], rec.get("alias")) if rec["_row"] != PENDING else rec["alias"]
                        if cur is None or cur != want:  # needRecordsUpdate (:373-381)
                            r53_ops.append((head(9, 1, kind), row, (j << 20) | k, z["_row"], acc["_row"], rec["_row"]))
                            if rec["_row"] == PENDING:
                                rec["alias"] = want
                            else:
                                new_alias[rec["_row"]] = want
                if created:
                    ev |= EV_CREATED
            return status(ST_OK, 0, ev)

        st_r53.append(r53())

    cache = {(0 if _resource(ob) == "service" else 1, ob.get("ns", "default"), ob["name"]) for ob in objects}

    def parse_owner(s):
        p = s.split("/")
        if len(p) != 3 or p[0] not in ("service", "ingress"):
            return None
        return (0 if p[0] == "service" else 1, p[1], p[2])

    ga_orph = []
    for acc in accs:
        tags = {}
        for k, v in acc.get("tags", []):
            tags[k] = v
        if tags.get(TAG_MANAGED, "") != "true" or tags.get(TAG_CLUSTER, "") != cluster:
            continue
        key = parse_owner(tags.get(TAG_OWNER, ""))
        if key is None or key in cache:
            continue
        delete_chain(ga_orph, NONE, 0, acc)

    r53_orph = []
    prefix = f'"heritage=aws-global-accelerator-controller,cluster={cluster},'

    def orphan(v):
        if not v.startswith(prefix) or len(v) < len(prefix) + 1 or not v.endswith('"'):
            return False
        key = parse_owner(v[len(prefix):-1])
        return key is not None and key not in cache

    h10 = head(10, 1, 0)
    for z in zones:
        ov = [(vi, v, r) for r in z["_recs"] for vi, v in r["_vals"] if orphan(v)]
        for r in z["_recs"]:
            if r.get("alias") is None:
                continue
            seen = set()
            for vi, v, vr in ov:
                if vr["name"] != r["name"] or v in seen:
                    continue
                seen.add(v)
                r53_orph.append((h10, NONE, 0, z["_row"], r["_row"], vi))
        for vi, v, vr in ov:
            r53_orph.append((h10, NONE, 1, z["_row"], vr["_row"], vi))

    ops = ga_ops + ga_orph + r53_ops + r53_orph
    sb = [0, len(ga_ops), len(ga_ops) + len(ga_orph), len(ga_ops) + len(ga_orph) + len(r53_ops), len(ops)]
    return dict(status_ga=st_ga, status_r53=st_r53, derived=derived, ops=ops, section_begin=sb, tok=tok, dports=dports)


# ---------------------------------------------------------------- EndpointGroupBinding (pkg/controller/endpointgroupbinding/reconcile.go)
#
# Second, independently written restatement (the C++ one is oracle.cpp orc_bindings_diff).  Go slices are modelled as
# (backing list, offset, length, capacity) so that `append(s[:i], s[i+1:]...)` aliases exactly as it does in Go.

class _GoPanic(Exception):
    pass


class _Slice:
    def __init__(self, backing, off, ln, cap):
        self.b, self.off, self.len, self.cap = backing, off, ln, cap

    def get(self, i):
        if not 0 <= i < self.len:
            raise _GoPanic("index out of range")
        return self.b[self.off + i]

    def sub(self, lo, hi=None):  # s[lo:hi]; hi defaults to len, may reach cap
        hi = self.len if hi is None else hi
        if not (0 <= lo <= hi <= self.cap):
            raise _GoPanic("slice bounds out of range")
        return _Slice(self.b, self.off + lo, hi - lo, self.cap - lo)

    def append_all(self, other):  # append(self, other...) without reallocation when capacity suffices
        vals = [other.get(i) for i in range(other.len)]
        if self.len + len(vals) <= self.cap:
            for k, v in enumerate(vals):
                self.b[self.off + self.len + k] = v
            return _Slice(self.b, self.off, self.len + len(vals), self.cap)
        nb = [self.b[self.off + i] for i in range(self.len)] + vals
        return _Slice(nb, 0, len(nb), len(nb))


EGB_ADD_FINALIZER, EGB_REMOVE_FINALIZER, EGB_REMOVE_ENDPOINT, EGB_ADD_ENDPOINT, EGB_UPDATE_WEIGHT, EGB_UPDATE_STATUS = range(11, 17)
ST_OK, ST_REQUEUE_30S, ST_ERR_RETRY, ST_PANIC, ST_REQUEUE_1S = 1, 3, 5, 7, 8
D_LB_NOT_FOUND, D_REF_NOT_FOUND, D_EG_NOT_FOUND = 5, 12, 13


def bindings_diff(objects, actual, bindings, known_egs):
    """-> (status words, ops as (head, binding, 0, a0, NONE, NONE)); rows as tables.pack / pack_bindings number them."""
    lbs = (actual or {}).get("lbs", [])
    for i, lb in enumerate(lbs):
        lb["_row"] = i
    cache = {(ob.get("kind", "service"), ob.get("ns", "default"), ob["name"]): ob for ob in reversed(objects)}  # lister: first row wins

    def get_lb(region, name):  # load_balancer.go:13-30 on the client of `region`
        for lb in lbs:
            if lb["region"] == region and lb["name"] == name:
                return lb
        return None

    statuses, ops = [], []
    ep_base = 0
    for k, b in enumerate(bindings):
        ids = list(b.get("endpoint_ids", []))
        rows = list(range(ep_base, ep_base + len(ids)))  # global ep_id rows
        ep_base += len(ids)

        def put(code, a0=NONE):
            ops.append((head(code, 2, 0), k, 0, a0, NONE, NONE))

        def reconcile():
            if b.get("deleting"):  # reconcileDelete :35-96
                if len(ids) == 0 or b.get("eg_arn", "") not in known_egs:
                    put(EGB_REMOVE_FINALIZER)
                    return status(ST_OK)
                status_ids = _Slice(list(rows), 0, len(rows), len(rows))  # obj.Status.EndpointIds
                endpoint_ids = status_ids.sub(0)                           # endpointIds := obj.Status.EndpointIds
                try:
                    for i in range(status_ids.len):                        # len evaluated once by `range`
                        put(EGB_REMOVE_ENDPOINT, status_ids.get(i))
                        endpoint_ids = endpoint_ids.sub(0, i).append_all(endpoint_ids.sub(i + 1))
                except _GoPanic:
                    return status(ST_PANIC)
                put(EGB_UPDATE_STATUS)
                return status(ST_REQUEUE_1S)
            if not b.get("finalizers", True):  # reconcileCreate :98-110
                put(EGB_ADD_FINALIZER)
                return status(ST_OK)
            # reconcileUpdate :112-217
            hostnames = []
            ref = b.get("ref")
            if ref is not None:
                ob = cache.get((ref[0], b.get("ns", "default"), ref[1]))
                if ob is None:
                    return status(ST_ERR_RETRY, D_REF_NOT_FOUND)
                hostnames = list(ob.get("lb_ingress", []))
            arns, order, first_row, regional = {}, [], {}, None
            for h in hostnames:
                code, name, region = _get_lb_name(h)
                if code >= 3:
                    return status(ST_ERR_RETRY, code - 4)  # NOT_ELB 5 -> detail 1, ... NLB 8 -> detail 4
                regional = region
                lb = get_lb(region, name)
                if lb is None:
                    return status(ST_ERR_RETRY, D_LB_NOT_FOUND)
                if lb["arn"] not in arns:
                    order.append(lb["arn"])
                    first_row[lb["arn"]] = lb["_row"]
                arns[lb["arn"]] = name
            new_ids = [a for a in order if a not in ids]
            removed = [r for r, e in zip(rows, ids) if e not in arns]
            if not new_ids and not removed and b.get("observed", True):
                return status(ST_OK)
            if b.get("eg_arn", "") not in known_egs:
                return status(ST_ERR_RETRY, D_EG_NOT_FOUND)
            for r in removed:
                if regional is None:
                    return status(ST_PANIC)  # nil *cloudaws.AWS
                put(EGB_REMOVE_ENDPOINT, r)
            for a in new_ids:
                lb = get_lb(regional, arns[a])
                if lb is None:
                    return status(ST_ERR_RETRY, D_LB_NOT_FOUND)
                if lb.get("state", "active") != "active":
                    return status(ST_REQUEUE_30S)
                put(EGB_ADD_ENDPOINT, lb["_row"])
            for a in order:
                put(EGB_UPDATE_WEIGHT, first_row[a])
            put(EGB_UPDATE_STATUS)
            return status(ST_OK)

        statuses.append(reconcile())
    return statuses, ops


def _get_lb_name(hostname):
    """GetLBNameFromHostname WITHOUT DetectCloudProvider in front (the binding controller calls it directly, reconcile.go:125)."""
    if _ALB.search(hostname) or _NLB.search(hostname):
        return tokenise(hostname)
    return (5, None, None)
