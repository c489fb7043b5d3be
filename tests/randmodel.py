"""Randomized small snapshots (dict model of tables.pack) that hit every branch of the path:
drifted / missing / duplicate accelerators, listener and endpoint-group cardinalities, bad hostnames,
unmanaged cleanup, orphans, wildcard and duplicate Route53 records, zone walks, odd annotations."""
import random

ANN = "aws-global-accelerator-controller.h3poteto.dev/"
REGIONS = ["us-east-1", "ap-northeast-1", "eu-west-1"]
ZONES = ["example.com.", "h3poteto-test.dev.", "sub.example.com.", "dev.", "other.org."]
LISTEN_PORTS = [
    '[{"HTTP": 80}, {"HTTPS": 443}]', '[{"HTTPS":443}]', '[{"http": 8080}]', '[]', 'null', '[{"HTTP": 80}', '{"HTTP":80}',
    '[{"HTTP": "80"}]', '[{"HTTP": 80, "HTTPS": 443, "x": [1,2,{"y":null}]}]', '[{"HTTP": 80.5}]', ' [ {"HTTPS" : 8443 } , null ] ',
    '[{"HTTP": 80, "HTTP": 81}]', '[{"HTTPS": 4294967739}]',
]


def _hex(rng, n):
    return "".join(rng.choice("0123456789abcdef") for _ in range(n))


def lb_hostname(rng, name, region, kind):
    if kind == "nlb":
        return f"{name}-{_hex(rng, 16)}.elb.{region}.amazonaws.com"
    if kind == "alb":
        return f"{name}-{rng.randrange(10**8, 10**10)}.{region}.elb.amazonaws.com"
    return f"internal-{name}-{rng.randrange(10**8, 10**10)}.{region}.elb.amazonaws.com"


BAD_HOSTNAMES = ["example.com", "localhost", "", "foo.example.amazonaws.com", "abc-.elb.us-east-1.amazonaws.com",
                 "internal-x.us-east-1.elb.amazonaws.com", "nohyphen.us-east-1.elb.amazonaws.com", "x-1.elb.amazonaws.com",
                 "a.b", "1.2.3.4"]


def make(seed: int, n_objects: int = 24, cluster: str = "default"):
    rng = random.Random(seed)
    objects, lbs, accs, zones = [], [], [], [{"id": f"/hostedzone/Z{i}", "name": z, "records": []} for i, z in enumerate(ZONES)]
    if rng.random() < 0.3:  # duplicate zone name: first row must win
        zones.append({"id": "/hostedzone/ZDUP", "name": "example.com.", "records": []})
    acc_id = [0]

    def new_acc(owner, hostname, name, ports, proto="TCP", lb_arn=None, clus=cluster, managed="true", extra_tags=(), enabled=True, n_lis=1, n_eg=1):
        acc_id[0] += 1
        i = acc_id[0]
        tags = [("aws-global-accelerator-controller-managed", managed), ("aws-global-accelerator-owner", owner),
                ("aws-global-accelerator-target-hostname", hostname), ("aws-global-accelerator-cluster", clus)] + list(extra_tags)
        if rng.random() < 0.2:
            rng.shuffle(tags)
        lis = []
        for li in range(n_lis):
            egs = [{"arn": f"arn:aws:globalaccelerator::1:accelerator/{i}/listener/{li}/endpoint-group/{e}", "endpoints": [lb_arn] if lb_arn else []} for e in range(n_eg)]
            lis.append({"arn": f"arn:aws:globalaccelerator::1:accelerator/{i}/listener/{li}", "proto": proto, "ports": list(ports), "egs": egs})
        a = {"arn": f"arn:aws:globalaccelerator::1:accelerator/{i}", "name": name, "dns": f"a{i:08x}.awsglobalaccelerator.com", "enabled": enabled, "tags": tags, "listeners": lis}
        accs.append(a)
        return a

    for i in range(n_objects):
        kind = rng.choice(["service", "ingress"])
        ns = rng.choice(["default", "prod", "kube-system"])
        name = f"{kind[:3]}-{i}"
        ann = {}
        ob = dict(kind=kind, ns=ns, name=name, annotations=ann)
        region = rng.choice(REGIONS)
        lbname = f"k8s-{ns}-{name}-{_hex(rng, 10)}" if kind == "ingress" else _hex(rng, 32)
        lbkind = rng.choice(["alb", "alb-int"]) if kind == "ingress" else "nlb"
        host = lb_hostname(rng, lbname, region, lbkind)
        r = rng.random()
        if kind == "service":
            ob["spec_type"] = "LoadBalancer" if r > 0.08 else rng.choice(["ClusterIP", "NodePort"])
            if rng.random() < 0.85:
                ann["service.beta.kubernetes.io/aws-load-balancer-type"] = rng.choice(["nlb", "external"])
            elif rng.random() < 0.5:
                ob["lb_class"] = True
            nports = rng.choice([1, 2, 2, 3, 5])
            ob["ports"] = [(rng.choice([80, 443, 8080, 8443, 53, 9000]), rng.choice(["TCP", "TCP", "TCP", "UDP", "udp", "Tcp", "SCTP"])) for _ in range(nports)]
        else:
            c = rng.random()
            if c < 0.7:
                ob["ingress_class"] = "alb"
            elif c < 0.8:
                ob["ingress_class"] = "nginx"
            elif c < 0.9:
                ann["kubernetes.io/ingress.class"] = rng.choice(["alb", "nginx"])
            ob["ports"] = [rng.choice([80, 8080, 0, 443]) for _ in range(rng.choice([0, 1, 1, 2, 3]))]
            if rng.random() < 0.6:
                ann["alb.ingress.kubernetes.io/listen-ports"] = rng.choice(LISTEN_PORTS)
        managed = rng.random() < 0.8
        if managed:
            ann[ANN + "global-accelerator-managed"] = rng.choice(["true", "yes", ""])
        if rng.random() < 0.2:
            ann[ANN + "global-accelerator-name"] = rng.choice(["custom-name", ""])
        if rng.random() < 0.25:
            ann[ANN + "global-accelerator-tags"] = rng.choice(["Environment=foo,Service=bar", "a=b", "bad", "a=b=c,k=v", "=v,k=", "k=v,k=w", "aws-global-accelerator-owner=evil"])
        if rng.random() < 0.2:
            ann[ANN + "client-ip-preservation"] = rng.choice(["true", "false"])
        if rng.random() < 0.2:
            ann[ANN + "ip-address-type"] = rng.choice(["ipv4", "IPV4", "dualstack", "bogus"])
        for _ in range(rng.randrange(0, 4)):
            ann[f"example.com/noise-{_hex(rng, 4)}"] = _hex(rng, rng.randrange(0, 20))
        # lbIngress
        hosts = []
        r = rng.random()
        if r < 0.06:
            pass
        elif r < 0.16:
            hosts = [rng.choice(BAD_HOSTNAMES)]
        elif r < 0.22:
            hosts = [rng.choice(BAD_HOSTNAMES[:2] + BAD_HOSTNAMES[8:]), host] if rng.random() < 0.5 else [host, lb_hostname(rng, lbname + "b", region, lbkind)]
        else:
            hosts = [host]
        ob["lb_ingress"] = hosts
        # actual LBs
        lb_arns = {}
        for h in hosts:
            if ".elb." not in h or rng.random() < 0.06:
                continue
            from_h = h.split(".")[0]
            nm = from_h.rsplit("-", 1)[0]
            if nm.startswith("internal-") and h.endswith(".elb.amazonaws.com"):
                nm = nm[len("internal-"):]
            reg = h.split(".")[1] if h.endswith(".elb.amazonaws.com") else h.split(".")[2]
            arn = f"arn:aws:elasticloadbalancing:{reg}:1:loadbalancer/net/{nm}/{_hex(rng, 16)}"
            state = "active" if rng.random() > 0.06 else rng.choice(["provisioning", "failed"])
            dns = h if rng.random() > 0.04 else "other-" + h
            lbs.append({"region": reg, "name": nm, "dns": dns, "arn": arn, "state": state})
            if rng.random() < 0.05:
                lbs.append({"region": reg, "name": nm, "dns": "dup-" + h, "arn": arn + "dup", "state": "active"})
            lb_arns[h] = arn
        # actual accelerators owned by this object
        owner = f"{kind}/{ns}/{name}"
        dname = ann.get(ANN + "global-accelerator-name") or f"{kind}-{ns}-{name}"
        dports = [p[0] for p in ob.get("ports", [])] if kind == "service" else list(ob.get("ports", []))
        proto = "TCP"
        if kind == "service":
            for _, p in ob["ports"]:
                if p.lower() == "udp":
                    proto = "UDP"
                elif p.lower() == "tcp":
                    proto = "TCP"
        r = rng.random()
        h0 = hosts[0] if hosts else host
        arn0 = lb_arns.get(h0)
        utags = [t.split("=") for t in ann.get(ANN + "global-accelerator-tags", "").split(",") if len(t.split("=")) == 2]
        if r < 0.12:
            pass  # missing -> create
        else:
            kw = dict(owner=owner, hostname=h0, name=dname, ports=dports, proto=proto, lb_arn=arn0, extra_tags=[tuple(t) for t in utags])
            d = rng.random()
            if d < 0.55:
                pass
            elif d < 0.62:
                kw["ports"] = dports + [9999]
            elif d < 0.67:
                kw["ports"] = dports[:-1]
            elif d < 0.72:
                kw["proto"] = "UDP" if proto == "TCP" else "TCP"
            elif d < 0.76:
                kw["name"] = "stale-name"
            elif d < 0.80:
                kw["extra_tags"] = []
            elif d < 0.83:
                kw["enabled"] = False
            elif d < 0.87:
                kw["n_lis"] = rng.choice([0, 2])
            elif d < 0.91:
                kw["n_eg"] = rng.choice([0, 2])
            elif d < 0.94:
                kw["lb_arn"] = "arn:aws:elasticloadbalancing:other"
            elif d < 0.96:
                kw["hostname"] = "stale." + h0
            elif d < 0.98:
                kw["clus"] = "other-cluster"
            else:
                kw["managed"] = "false"
            new_acc(**kw)
            if rng.random() < 0.06:
                new_acc(**kw)  # duplicate accelerator for the same owner: both are updated
        # route53
        r53names = []
        if rng.random() < 0.7:
            nh = rng.choice([1, 1, 2, 3])
            for k in range(nh):
                z = rng.choice(ZONES[:4])[:-1]
                c = rng.random()
                if c < 0.1:
                    r53names.append(f"*.w{i}.{z}")
                elif c < 0.15:
                    r53names.append(f"h{i}-{k}.nozone.invalid")
                elif c < 0.18:
                    r53names.append(rng.choice(["", " sp.example.com", "example.com"]))
                else:
                    r53names.append(f"h{i}-{k}.{rng.choice(['', 'a.', 'a.b.'])}{z}")
            ann[ANN + "route53-hostname"] = ",".join(r53names)
        ov = f'"heritage=aws-global-accelerator-controller,cluster={cluster},{kind}/{ns}/{name}"'
        owned = list(r53names)
        if not r53names and rng.random() < 0.4:
            owned = [f"old{i}.example.com"]  # annotation removed -> cleanup deletes these
        acc_dns = accs[-1]["dns"] if accs else "none.awsglobalaccelerator.com"
        for hn in owned:
            if rng.random() < 0.2:
                continue  # missing -> create
            zn = None
            t = hn
            while t:
                if t + "." in ZONES:
                    zn = t + "."
                    break
                t = ".".join(t.split(".")[1:])
            if zn is None:
                continue
            zone = next(z for z in zones if z["name"] == zn)
            rn = hn.replace("*", "\\052", 1) + "."
            alias = acc_dns + "." if rng.random() > 0.15 else "stale.awsglobalaccelerator.com."
            recs = [{"name": rn, "type": "TXT", "values": [ov] if rng.random() > 0.1 else [ov, ov]},
                    {"name": rn, "type": "A", "alias": alias}]
            if rng.random() < 0.15:
                recs.insert(1, {"name": rn, "type": "AAAA", "alias": alias})
            if rng.random() < 0.1:
                recs[-1] = {"name": rn, "type": "A", "values": ["1.2.3.4"]}  # not an alias
            if rng.random() < 0.3:
                recs.reverse()
            zone["records"].extend(recs)
        objects.append(ob)

    # orphans: accelerators / records whose owner is not in the cache
    for k in range(rng.randrange(0, 4)):
        okind = rng.choice(["service", "ingress"])
        own = rng.choice([f"{okind}/default/gone-{k}", f"{okind}/default/gone-{k}/x", f"pod/default/gone-{k}", "", f"{okind}//gone-{k}"])
        new_acc(own, f"gone-{k}.elb.us-east-1.amazonaws.com", f"{okind}-default-gone-{k}", [80], lb_arn="arn:lb:gone",
                clus=rng.choice([cluster, cluster, "other"]), n_lis=rng.choice([0, 1, 1, 2]), n_eg=rng.choice([0, 1, 1, 2]))
        ovk = f'"heritage=aws-global-accelerator-controller,cluster={rng.choice([cluster, cluster, "zzz"])},{own}"'
        zone = rng.choice(zones)
        rn = f"gone{k}.{zone['name']}"
        zone["records"].append({"name": rn, "type": "TXT", "values": [ovk, '"unrelated"'] if rng.random() < 0.5 else [ovk]})
        zone["records"].append({"name": rn, "type": "A", "alias": "x.awsglobalaccelerator.com."})
        if rng.random() < 0.3:
            zone["records"].append({"name": rn, "type": "TXT", "values": [ovk]})
    # an accelerator claimed by two target-hostname matches (route53: count != 1)
    if accs and rng.random() < 0.5:
        src = rng.choice(accs)
        th = dict(src["tags"]).get("aws-global-accelerator-target-hostname", "x")
        new_acc("service/default/nobody", th, "dup-host", [80])
    for z in zones:
        for _ in range(rng.randrange(0, 3)):
            z["records"].append({"name": f"noise{_hex(rng, 4)}.{z['name']}", "type": rng.choice(["A", "CNAME", "TXT"]), "values": [_hex(rng, 8)]})
    if rng.random() < 0.5:
        rng.shuffle(accs)
    return objects, {"lbs": lbs, "accelerators": accs, "zones": zones}
