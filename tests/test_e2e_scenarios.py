"""The reference's own end-to-end expectations (local_e2e/e2e_test.go:90-221, helpers :257-385) as scenario tests of the batch
path: fixture object -> pack -> diff -> execute on the mock AWS model -> ... until nothing is left to do, then the predicates the
Go e2e suite polls for; delete the object -> everything it owned is gone.  The reference runs these against a real cluster and
a real AWS account; here the mock AWS of tests/executor.py stands in, and the same scenario runs through the oracle (which
these scenarios thereby pin at the level the reference itself tests), the device logic on hostsim, and the GPU.

Fixtures follow local_e2e/pkg/fixtures/service.go:10-51 (NewNLBService) and ingress.go:15-58 (NewALBIngress, port 443)."""
import importlib

import pytest

import executor

pyref = importlib.import_module("oracle.pyref")
ANN = "aws-global-accelerator-controller.h3poteto.dev/"
CLUSTER = "e2e"
HOSTNAME = "foo.h3poteto-test.dev,bar.h3poteto-test.dev"  # E2E_HOSTNAME may hold several names (e2e_test.go:101)
NLB_HOST = "a1b2c3d4e5f6a7b8c9d0e1f2a3b4c5d6-0123456789abcdef.elb.ap-northeast-1.amazonaws.com"
ALB_HOST = "k8s-default-e2etest-0a1b2c3d4e-1234567890.ap-northeast-1.elb.amazonaws.com"


def nlb_service(ns="default", name="e2e-test", hostname=HOSTNAME):  # fixtures.NewNLBService
    return dict(kind="service", ns=ns, name=name, spec_type="LoadBalancer",
                annotations={ANN + "global-accelerator-managed": "true", ANN + "route53-hostname": hostname,
                             "service.beta.kubernetes.io/aws-load-balancer-backend-protocol": "tcp",
                             "service.beta.kubernetes.io/aws-load-balancer-cross-zone-load-balancing-enabled": "true",
                             "service.beta.kubernetes.io/aws-load-balancer-type": "nlb",
                             "service.beta.kubernetes.io/aws-load-balancer-scheme": "internet-facing"},
                ports=[(80, "TCP"), (443, "TCP")], lb_ingress=[NLB_HOST])


def alb_ingress(ns="default", name="e2e-test", hostname=HOSTNAME, port=443):  # fixtures.NewALBIngress
    return dict(kind="ingress", ns=ns, name=name, ingress_class="alb",
                annotations={ANN + "global-accelerator-managed": "true", ANN + "route53-hostname": hostname,
                             "alb.ingress.kubernetes.io/scheme": "internet-facing", "alb.ingress.kubernetes.io/certificate-arn": "arn:aws:acm:x",
                             "alb.ingress.kubernetes.io/listen-ports": '[{"HTTPS":%d}]' % port},
                ports=[80], lb_ingress=[ALB_HOST])


def aws_account():
    """What the e2e environment provides before the controller acts: the load balancers the cloud controllers made, the hosted zone."""
    return {"lbs": [{"region": "ap-northeast-1", "name": "a1b2c3d4e5f6a7b8c9d0e1f2a3b4c5d6", "dns": NLB_HOST, "arn": "arn:aws:elasticloadbalancing:ap-northeast-1:1:loadbalancer/net/a1b2/x", "state": "active"},
                    {"region": "ap-northeast-1", "name": "k8s-default-e2etest-0a1b2c3d4e", "dns": ALB_HOST, "arn": "arn:aws:elasticloadbalancing:ap-northeast-1:1:loadbalancer/app/k8s/y", "state": "active"}],
            "accelerators": [], "zones": [{"id": "/hostedzone/Z1", "name": "h3poteto-test.dev.", "records": [{"name": "h3poteto-test.dev.", "type": "NS", "values": ["ns-1.awsdns.com."]}]}]}


# ---- the reference's helpers, over the dict model (same list / filter logic as the SDK wrappers they call)

def list_by_resource(actual, resource, ob):  # ListGlobalAcceleratorByResource (global_accelerator.go:87-110)
    t = {pyref.TAG_MANAGED: "true", pyref.TAG_OWNER: f"{resource}/{ob['ns']}/{ob['name']}", pyref.TAG_CLUSTER: CLUSTER}
    return [a for a in actual["accelerators"] if pyref.tags_contain(a.get("tags", []), t)]


def global_accelerator_is_created(actual, lb, resource, ob):  # waitUntilGlobalAccelerator (e2e_test.go:257-304)
    for acc in list_by_resource(actual, resource, ob):
        if len(acc.get("listeners", [])) != 1:  # GetListener: exactly one
            return False
        egs = acc["listeners"][0].get("egs", [])
        if len(egs) != 1:
            return False
        if lb["arn"] in egs[0].get("endpoints", []):
            return True
    return False


def owned_alias_sets(zone, ov):  # FindOwneredARecordSets (route53.go:216-238)
    names = [r["name"] for r in zone["records"] if ov in r.get("values", [])]
    return [r for r in zone["records"] if r.get("alias") is not None and r["name"] in names]


def route53_is_created(actual, hostnames, lb_hostname, resource, ob):  # waitUntilRoute53 (e2e_test.go:306-343)
    t = {pyref.TAG_MANAGED: "true", pyref.TAG_HOST: lb_hostname, pyref.TAG_CLUSTER: CLUSTER}
    accs = [a for a in actual["accelerators"] if pyref.tags_contain(a.get("tags", []), t)]
    if not accs:
        return False
    ov = pyref.owner_value(CLUSTER, resource, ob["ns"], ob["name"])
    zone = actual["zones"][0]
    for _h in hostnames:
        if not any(r["alias"] == accs[0]["dns"] + "." for r in owned_alias_sets(zone, ov)):
            return False
    return True


def everything_is_cleaned_up(actual, resource, ob):  # waitUntilCleanup (e2e_test.go:345-385)
    ov = pyref.owner_value(CLUSTER, resource, ob["ns"], ob["name"])
    return not owned_alias_sets(actual["zones"][0], ov) and not list_by_resource(actual, resource, ob)


def converge(garecon, diff_fn, objects, actual, max_rounds=6):
    for _ in range(max_rounds):
        cs = diff_fn(garecon.pack(objects, actual))
        if len(cs.ops) == 0:
            return actual
        actual = executor.apply(objects, actual, cs, cluster=CLUSTER)
    raise AssertionError("the controller did not settle")


def scenario(garecon, diff_fn, ob, lb_index, allow_empty_cache_diff_fn=None):
    resource = "service" if ob["kind"] == "service" else "ingress"
    actual = aws_account()
    lb = actual["lbs"][lb_index]
    hostnames = HOSTNAME.split(",")
    # "Resources should be created"
    actual = converge(garecon, diff_fn, [ob], actual)
    assert global_accelerator_is_created(actual, lb, resource, ob)
    assert len(list_by_resource(actual, resource, ob)) == 1
    assert route53_is_created(actual, hostnames, ob["lb_ingress"][0], resource, ob)
    ov = pyref.owner_value(CLUSTER, resource, ob["ns"], ob["name"])
    for h in hostnames:  # one TXT owner record and one alias A record per hostname (route53.go:100-113)
        sets = [r for r in actual["zones"][0]["records"] if r["name"] == h + "."]
        assert sorted(r["type"] for r in sets) == ["A", "TXT"]
        assert [r for r in sets if r["type"] == "TXT"][0]["values"] == [ov]
    created = actual
    # "Remove resources": the object leaves the cache; its resources are the orphan sections of the next diff
    actual = converge(garecon, allow_empty_cache_diff_fn or diff_fn, [], actual)
    assert everything_is_cleaned_up(actual, resource, ob)
    assert actual["accelerators"] == [] and [r["type"] for r in actual["zones"][0]["records"]] == ["NS"]  # nothing else was touched
    return created


def oracle_diff(oracle):
    return lambda snap: oracle.diff(snap, CLUSTER, mode=1)


def test_service_scenario_on_the_oracle(garecon, oracle):
    created = scenario(garecon, oracle_diff(oracle), nlb_service(), 0)
    acc = created["accelerators"][0]
    assert acc["listeners"][0]["ports"] == [80, 443] and acc["listeners"][0]["proto"] == "TCP"  # listenerForService (:503-515)


def test_ingress_scenario_on_the_oracle(garecon, oracle):
    """e2e_test.go:192-205 "Check Listener ports": exactly one accelerator, one port range, 443-443."""
    created = scenario(garecon, oracle_diff(oracle), alb_ingress(), 1)
    accs = list_by_resource(created, "ingress", alb_ingress())
    assert len(accs) == 1
    assert accs[0]["listeners"][0]["ports"] == [443]


def test_faithful_mode_agrees_on_every_round(garecon, oracle):
    for ob, k in ((nlb_service(), 0), (alb_ingress(), 1)):
        def diff(snap):
            a, b, c = oracle.diff(snap, CLUSTER, mode=0), oracle.diff(snap, CLUSTER, mode=1), oracle.diff(snap, CLUSTER, mode=2, threads=2)
            assert a.diff(b) == [] and a.diff(c) == []
            return a
        scenario(garecon, diff, ob, k)


@pytest.fixture(scope="module")
def hostsim_engines(garecon):
    import __graft_entry__ as ge
    lib = garecon.abi.load_library(ge.build_hostsim())
    a, b = garecon.Engine(cluster_name=CLUSTER, lib=lib), garecon.Engine(cluster_name=CLUSTER, lib=lib, allow_empty_cache=True)
    yield a, b
    a.close()
    b.close()


def _engine_diff(e, keep):
    def diff(snap):
        keep.append(snap)  # hostsim reads the columns in place
        e.load(snap)
        return e.diff()
    return diff


@pytest.mark.parametrize("which", ["service", "ingress"])
def test_scenarios_on_the_device_logic(garecon, hostsim_engines, which):
    """The last object's deletion leaves an EMPTY cache: the default engine refuses to sweep then (garecon.h "Orphan sweep
    precondition"), the worker that knows its informer is synced uses GAR_FLAG_ALLOW_EMPTY_CACHE."""
    e, e_empty = hostsim_engines
    keep = []
    ob, k = (nlb_service(), 0) if which == "service" else (alb_ingress(), 1)
    scenario(garecon, _engine_diff(e, keep), ob, k, allow_empty_cache_diff_fn=_engine_diff(e_empty, keep))
    snap = garecon.pack([], executor.apply([ob], aws_account(), _engine_diff(e, keep)(garecon.pack([ob], aws_account())), cluster=CLUSTER))
    keep.append(snap)
    e.load(snap)
    with pytest.raises(garecon.abi.GarError, match="object table is empty"):
        e.diff()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["service", "ingress"])
def test_scenarios_on_the_gpu(garecon, oracle, which):
    ob, k = (nlb_service(), 0) if which == "service" else (alb_ingress(), 1)
    with garecon.Engine(cluster_name=CLUSTER) as e, garecon.Engine(cluster_name=CLUSTER, allow_empty_cache=True) as e_empty:
        def mk(eng):
            def diff(snap):
                eng.load(snap)
                got = eng.diff()
                want = oracle.diff(snap, CLUSTER, mode=1)
                assert got.diff(want) == [], got.describe_first_mismatch(want)
                return got
            return diff
        scenario(garecon, mk(e), ob, k, allow_empty_cache_diff_fn=mk(e_empty))
