"""Adversarial snapshot: keys duplicated far beyond the per-bucket fast path of the index build (forces the stable
radix-sort fallback) and beyond the per-object op staging / owned-row caches (forces the re-evaluation paths)."""
ANN = "aws-global-accelerator-controller.h3poteto.dev/"


def make(ndup_acc=70, ndup_alias=60, ndup_val=80):
    host = "aaaabbbbccccddddaaaabbbbccccdddd-1234567890abcdef.elb.us-east-1.amazonaws.com"
    lbarn = "arn:aws:elasticloadbalancing:us-east-1:1:loadbalancer/net/aaaabbbbccccddddaaaabbbbccccdddd/1"
    hot = dict(kind="service", ns="default", name="hot", annotations={ANN + "global-accelerator-managed": "true", ANN + "route53-hostname": "a.example.com,b.example.com",
                                                                      "service.beta.kubernetes.io/aws-load-balancer-type": "nlb"},
               lb_ingress=[host], ports=[(80, "TCP"), (443, "TCP")])
    gone = dict(kind="ingress", ns="default", name="unannotated", ingress_class="alb", annotations={}, lb_ingress=[host], ports=[80])
    objects = [hot, gone] + [dict(kind="service", ns="default", name=f"s{i}", annotations={"service.beta.kubernetes.io/aws-load-balancer-type": "nlb"}, lb_ingress=[], ports=[(80, "TCP")]) for i in range(40)]
    accs = []
    for k in range(ndup_acc):  # every one of them is updated, in list order
        tags = [("aws-global-accelerator-controller-managed", "true"), ("aws-global-accelerator-owner", "service/default/hot"),
                ("aws-global-accelerator-target-hostname", host if k == 0 else f"other{k}." + host), ("aws-global-accelerator-cluster", "default")]
        accs.append(dict(arn=f"acc{k}", name="service-default-hot" if k % 3 else "stale", dns=f"a{k}.awsglobalaccelerator.com", enabled=True, tags=tags,
                         listeners=[dict(arn=f"l{k}", proto="TCP", ports=[80, 443] if k % 2 else [80], egs=[dict(arn=f"e{k}", endpoints=[lbarn] if k % 5 else [])])]))
    ov_hot = '"heritage=aws-global-accelerator-controller,cluster=default,service/default/hot"'
    ov_un = '"heritage=aws-global-accelerator-controller,cluster=default,ingress/default/unannotated"'
    ov_orphan = '"heritage=aws-global-accelerator-controller,cluster=default,service/default/left-the-cache"'
    recs = [dict(name="a.example.com.", type="TXT", values=[ov_hot] * ndup_val)]
    recs += [dict(name="a.example.com.", type="AAAA" if k % 2 else "A", alias="stale.awsglobalaccelerator.com.") for k in range(ndup_alias)]
    recs += [dict(name="c.example.com.", type="TXT", values=[ov_un] * ndup_val + [ov_orphan] * ndup_val)]
    recs += [dict(name="c.example.com.", type="A", alias="x.awsglobalaccelerator.com.") for _ in range(ndup_alias)]
    actual = dict(lbs=[dict(region="us-east-1", name="aaaabbbbccccddddaaaabbbbccccdddd", dns=host, arn=lbarn, state="active")] * 20, accelerators=accs,
                  zones=[dict(id="Z1", name="example.com.", records=recs)])
    return objects, actual
