"""Dense multi-lbIngress / repeated-hostname models for the self-observation rules (include/garecon.h "Self-observation"):
objects with 2-3 lbIngress hostnames (distinct load balancers, repeated hostnames, two hostnames of one ARN), accelerators in
every starting state, user tags that overwrite the system tags, route53 annotations that repeat a hostname."""
import random

import randmodel

ANN = randmodel.ANN
TAGS = [None, "Environment=foo", "aws-global-accelerator-target-hostname=pinned.example.com", "aws-global-accelerator-controller-managed=false",
        "aws-global-accelerator-controller-managed=true", "aws-global-accelerator-owner=evil", "aws-global-accelerator-cluster=other",
        "aws-global-accelerator-cluster=default,a=b", "aws-global-accelerator-controller-managed=false,aws-global-accelerator-controller-managed=true"]


def make(seed: int, n_objects: int = 30, cluster: str = "default"):
    rng = random.Random(1000 + seed)
    objects, lbs, accs = [], [], []
    zones = [{"id": f"/hostedzone/Z{i}", "name": z, "records": []} for i, z in enumerate(randmodel.ZONES)]

    def acc(owner, thost, name, ports, proto, endpoints, extra=(), n_lis=1, n_eg=1, enabled=True):
        i = len(accs) + 1
        tags = [("aws-global-accelerator-controller-managed", "true"), ("aws-global-accelerator-owner", owner),
                ("aws-global-accelerator-target-hostname", thost), ("aws-global-accelerator-cluster", cluster)] + list(extra)
        lis = [{"arn": f"l{i}-{x}", "proto": proto, "ports": list(ports),
                "egs": [{"arn": f"e{i}-{x}-{y}", "endpoints": list(endpoints)} for y in range(n_eg)]} for x in range(n_lis)]
        a = {"arn": f"acc{i}", "name": name, "dns": f"a{i:06x}.awsglobalaccelerator.com", "enabled": enabled, "tags": tags, "listeners": lis}
        accs.append(a)
        return a

    for i in range(n_objects):
        kind = rng.choice(["service", "ingress"])
        ns, name = rng.choice(["default", "prod"]), f"{kind[:3]}-{i}"
        ann = {ANN + "global-accelerator-managed": "true"}
        ob = dict(kind=kind, ns=ns, name=name, annotations=ann)
        if kind == "service":
            ob["spec_type"] = "LoadBalancer"
            ann["service.beta.kubernetes.io/aws-load-balancer-type"] = "nlb"
            ob["ports"] = [(80, "TCP"), (443, "TCP")]
            dports, lbkind = [80, 443], "nlb"
        else:
            ob["ingress_class"] = "alb"
            ob["ports"] = [80]
            dports, lbkind = [80], rng.choice(["alb", "alb-int"])
        tags = rng.choice(TAGS)
        if tags is not None:
            ann[ANN + "global-accelerator-tags"] = tags
        nh = rng.choice([1, 2, 2, 3])
        hosts, arns = [], {}
        for j in range(nh):
            if hosts and rng.random() < 0.2:
                hosts.append(rng.choice(hosts))  # the same hostname again
                continue
            region = rng.choice(randmodel.REGIONS)
            lbname = f"k8s-{ns}-{name}-{j}{randmodel._hex(rng, 6)}" if kind == "ingress" else randmodel._hex(rng, 30) + f"{j:02d}"
            h = randmodel.lb_hostname(rng, lbname, region, lbkind)
            hosts.append(h)
            arn = f"arn:aws:elasticloadbalancing:{region}:1:loadbalancer/x/{lbname}"
            if arns and rng.random() < 0.15:
                arn = rng.choice(list(arns.values()))  # two hostnames, one ARN
            arns[h] = arn
            r = rng.random()
            if r < 0.05:
                continue  # load balancer missing: the object stops here
            lbs.append({"region": region, "name": lbname, "dns": h, "arn": arn, "state": "active" if r > 0.12 else "provisioning"})
        if rng.random() < 0.1:
            hosts.insert(rng.randrange(len(hosts) + 1), "not-aws.example.org")  # DetectCloudProvider error: `continue`
        ob["lb_ingress"] = hosts
        owner = f"{kind}/{ns}/{name}"
        dname = f"{kind}-{ns}-{name}"
        utags = [tuple(t.split("=")) for t in (tags or "").split(",") if len(t.split("=")) == 2]
        first = next((h for h in hosts if h in arns), None)
        state = rng.choice(["none", "sync", "sync", "both", "nolis", "noeg", "stale", "two", "manylis"])
        if first is not None and state != "none":
            eps = [arns[first]]
            if state == "both":
                eps = list(dict.fromkeys(arns.values()))
            kw = dict(owner=owner, thost=first, name=dname, ports=dports, proto="TCP", endpoints=eps, extra=utags if rng.random() < 0.7 else ())
            if state == "nolis":
                kw["n_lis"] = 0
            if state == "noeg":
                kw["n_eg"] = 0
            if state == "manylis":
                kw["n_lis"] = 2
            if state == "stale":
                kw.update(thost="stale." + first, endpoints=["arn:other"], name="old-name")
            acc(**kw)
            if state == "two":
                acc(**dict(kw, endpoints=["arn:other"]))
        # one accelerator per remaining load-balancer hostname so that the route53 path finds exactly one by hostname
        for h in dict.fromkeys(hosts):
            if h in arns and h != first and rng.random() < 0.8:
                acc(f"service/elsewhere/x{i}", h, f"other-{i}", [80], "TCP", [arns[h]])
        if rng.random() < 0.8:
            zn = rng.choice(randmodel.ZONES[:3])[:-1]
            names = [f"h{i}-{k}.{zn}" for k in range(rng.choice([1, 2, 3]))]
            if rng.random() < 0.5:
                names.insert(rng.randrange(len(names) + 1), rng.choice(names))  # a hostname repeated in the annotation
            if rng.random() < 0.1:
                names.append("*.w%d.%s" % (i, zn))
                names.append("*.w%d.%s" % (i, zn))
            ann[ANN + "route53-hostname"] = ",".join(names)
            ov = f'"heritage=aws-global-accelerator-controller,cluster={cluster},{kind}/{ns}/{name}"'
            zone = next(z for z in zones if z["name"] == zn + ".")
            for hn in dict.fromkeys(names):
                r = rng.random()
                if r < 0.4:
                    continue  # missing -> create
                alias = (accs[-1]["dns"] if accs else "x") + "." if r < 0.75 else "stale.awsglobalaccelerator.com."
                rn = hn.replace("*", "\\052", 1) + "."
                zone["records"] += [{"name": rn, "type": "TXT", "values": [ov]}, {"name": rn, "type": "A", "alias": alias}]
        objects.append(ob)
    return objects, {"lbs": lbs, "accelerators": accs, "zones": zones}
