"""A mock AWS backend for the dict model of tables.pack(): applies a change set the way the reference's SDK wrappers
would (global_accelerator.go:654-1013, route53.go:183-315), so tests can run diff -> apply -> diff to a fixed point.

This is the stand-in for the mock cloudprovider that BASELINE config 1 names but the reference does not have
(SURVEY.md fact 2), and a miniature of the Go-side executor of SURVEY.md §8 row f2.  Test infrastructure only.
"""
from __future__ import annotations

import copy
import importlib

pyref = importlib.import_module("oracle.pyref")

NONE = 0xFFFFFFFF
PENDING = 0xFFFFFFFE  # GAR_PENDING: the resource an earlier op of the same object created
ANN_IPPRESERVE = pyref.ANN_IPPRESERVE


def _flat(actual):
    """Flat row numbering identical to tables.pack / pyref._Rows."""
    accs, lis, egs, zones, recs, vals = [], [], [], [], [], []
    for a in actual.get("accelerators", []):
        accs.append(a)
        for li in a.get("listeners", []):
            lis.append((a, li))
            for eg in li.get("egs", []):
                egs.append((li, eg))
    for z in actual.get("zones", []):
        zones.append(z)
        for r in z.get("records", []):
            recs.append((z, r))
            for v in r.get("values", []):
                vals.append((r, v))
    return accs, lis, egs, zones, recs, vals


def apply(objects, actual, cs, cluster="default"):
    """Returns a NEW actual model with every op of change set `cs` (abi.ChangeSet) executed against `actual`."""
    actual = copy.deepcopy(actual)
    for k in ("accelerators", "lbs", "zones"):
        actual.setdefault(k, [])
    accs, lis, egs, zones, recs, _vals = _flat(actual)
    lbs = actual["lbs"]
    dead_accs, dead_recs = set(), set()
    serial = [len(accs)]
    last_created_listener = {}
    created_acc, pending_eg, created_rec = {}, {}, {}  # per object: what its own earlier ops created (GAR_PENDING arguments)

    def desired(ob):
        ports, proto, _ = pyref.desired_listener(ob)
        return list(ports), ("UDP" if proto == 1 else "TCP")

    def sys_tags(ob, lb, with_cluster):
        res = "service" if ob.get("kind", "service") == "service" else "ingress"
        t = [(pyref.TAG_MANAGED, "true"), (pyref.TAG_OWNER, f"{res}/{ob.get('ns', 'default')}/{ob['name']}"), (pyref.TAG_HOST, lb["dns"])]
        if with_cluster:
            t.append((pyref.TAG_CLUSTER, cluster))
        return t + list(pyref.accelerator_tags(ob))

    for op in cs.ops.tolist():
        head, obj, sub, a0, a1, a2 = (int(x) for x in op)
        code = head & 0xFF
        ob = objects[obj] if obj != NONE else None
        if code == 1:  # GA_CREATE_CHAIN: createAccelerator + createListener + createEndpointGroup (:213-252)
            lb = lbs[a0]
            serial[0] += 1
            ports, proto = desired(ob)
            eg = {"arn": f"e-new-{serial[0]}", "endpoints": [lb["arn"]]}
            new_acc = {
                "arn": f"arn:aws:globalaccelerator::1:accelerator/new-{serial[0]}", "name": pyref.accelerator_name(ob),
                "dns": f"new{serial[0]:06d}.awsglobalaccelerator.com", "enabled": True, "tags": sys_tags(ob, lb, True),
                "listeners": [{"arn": f"l-new-{serial[0]}", "proto": proto, "ports": ports, "egs": [eg]}]}
            actual["accelerators"].append(new_acc)
            created_acc[obj] = new_acc
            pending_eg[(obj, PENDING)] = eg
        elif code == 2:  # GA_UPDATE_ACCEL: UpdateAccelerator(Enabled, Name) + TagResource (:703-741)
            acc, lb = (created_acc[obj] if a0 == PENDING else accs[a0]), lbs[a1]
            acc["enabled"] = True
            acc["name"] = pyref.accelerator_name(ob)
            new = sys_tags(ob, lb, False)
            keys = {k for k, _ in new}
            final = {}
            for k, v in new:
                final[k] = v
            acc["tags"] = [(k, v) for k, v in acc.get("tags", []) if k not in keys] + list(final.items())
        elif code == 3:  # GA_CREATE_LISTENER (:815-837)
            acc = accs[a0]
            ports, proto = desired(ob)
            li = {"arn": f"l-new-{a0}-{len(acc.get('listeners', []))}", "proto": proto, "ports": ports, "egs": []}
            acc.setdefault("listeners", []).append(li)
            last_created_listener[a0] = li
        elif code == 4:  # GA_UPDATE_LISTENER (:839-861)
            _, li = lis[a1]
            li["ports"], li["proto"] = desired(ob)
        elif code == 5:  # GA_CREATE_EG (:971-990)
            li = last_created_listener[a0] if a1 == NONE else lis[a1][1]
            eg = {"arn": f"e-new-{a0}-{len(li.get('egs', []))}", "endpoints": [lbs[a2]["arn"]]}
            li.setdefault("egs", []).append(eg)
            pending_eg[(obj, a0)] = eg
        elif code == 6:  # GA_UPDATE_EG: EndpointConfigurations = [lb] (:992-1010)
            eg = pending_eg[(obj, a0)] if a1 == PENDING else egs[a1][1]
            eg["endpoints"] = [lbs[a2]["arn"]]
        elif code == 7:  # GA_DELETE_CHAIN (:254-272)
            dead_accs.add(a0)
        elif code == 8:  # R53_CREATE: TXT owner record, then A alias (:240-289)
            z, acc = zones[a0], accs[a1]
            k = sub & 0xFFFFF
            hn = dict(ob.get("annotations", {}))[pyref.ANN_R53].split(",")[k]
            res = "service" if ob.get("kind", "service") == "service" else "ingress"
            name = hn.replace("*", "\\052", 1) + "."
            z.setdefault("records", []).append({"name": name, "type": "TXT", "values": [pyref.owner_value(cluster, res, ob.get("ns", "default"), ob["name"])]})
            z["records"].append({"name": name, "type": "A", "alias": acc.get("dns", "") + "."})
            created_rec[(obj, hn)] = z["records"][-1]
        elif code == 9:  # R53_UPSERT_A (:291-315)
            if a2 == PENDING:
                hn = dict(ob.get("annotations", {}))[pyref.ANN_R53].split(",")[sub & 0xFFFFF]
                rec = created_rec[(obj, hn)]
            else:
                _, rec = recs[a2]
            rec["alias"] = accs[a1].get("dns", "") + "."
            rec["type"] = "A"
        elif code == 10:  # R53_DELETE_RECORD (:183-197)
            dead_recs.add(a1)
    dead_acc_ids = {id(accs[i]) for i in dead_accs}
    actual["accelerators"] = [a for a in actual["accelerators"] if id(a) not in dead_acc_ids]
    dead_rec_ids = {id(recs[i][1]) for i in dead_recs}
    for z in actual["zones"]:
        z["records"] = [r for r in z.get("records", []) if id(r) not in dead_rec_ids]
    return actual
