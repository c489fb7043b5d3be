"""Hand-derived vectors for the Global Accelerator ensure / update / cleanup sequencing, which the reference
pins only through the port / protocol predicate tables.  Every expectation is derived from the Go text:

  EnsureGlobalAcceleratorFor{Service,Ingress} (global_accelerator.go:112-211)
    * GetLoadBalancer miss -> error; DNS name != lbIngress hostname -> error; state != active -> Result{Requeue, 30s}; all before
      anything is listed or touched;
    * no accelerator for (cluster, owner) -> create the chain (created = true -> GlobalAcceleratorCreated event);
    * else EVERY listed accelerator is updated, in list order; the first error ends the object;
  updateGlobalAcceleratorFor* (:290-410), per accelerator, in this order:
    acceleratorChanged (disabled / name / tags) -> updateAccelerator;
    GetListener: 0 listeners -> createListener, > 1 -> error "Too many listeners" (:808-810);
    protocol or port predicate -> updateListener (never on a listener just created from the desired state);
    GetEndpointGroup: 0 -> createEndpointGroup, > 1 -> error "Too many endpoint groups" (:902-904);
    !endpointContainsLB -> updateEndpointGroup (never on a group just created with the LB);
  no managed annotation (globalaccelerator/service.go:64-84): CleanupGlobalAccelerator of every accelerator of the owner +
    GlobalAcceleratorDeleted event, every time — after the "no lbIngress" early return (:59-62);
  CleanupGlobalAccelerator / listRelatedGlobalAccelerator (:254-288): the endpoint group is deleted only when GetEndpointGroup
    succeeds (exactly one), the listener only when GetListener succeeds (exactly one), the accelerator always.

One Service `default/s` (port 80/TCP) behind an active NLB; a vector varies the object, the LB row and the accelerators; the
observable is the GA status word and the GA ops of the object."""
import pytest

ANN = "aws-global-accelerator-controller.h3poteto.dev/"
NAME32 = "0123456789abcdef0123456789abcdef"
HOST = NAME32 + "-0123456789abcdef.elb.us-west-2.amazonaws.com"
ALB_HOST = "k8s-default-i-0123456789-1234567890.us-west-2.elb.amazonaws.com"
NONE = 0xFFFFFFFF

OK, SKIP_NO_LB, REQ30, RETRY = 1, 2, 3, 5
D_LB_NOT_FOUND, D_LB_DNS, D_MANY_LIS, D_MANY_EGS = 5, 6, 7, 8
EV_CREATED, EV_DELETED = 1, 2
CREATE_CHAIN, UPDATE_ACCEL, CREATE_LIS, UPDATE_LIS, CREATE_EG, UPDATE_EG, DELETE_CHAIN = 1, 2, 3, 4, 5, 6, 7


def lb(dns=HOST, state="active", name=NAME32):
    return {"region": "us-west-2", "name": name, "dns": dns, "arn": "arn:lb", "state": state}


def eg(i, endpoints=("arn:lb",)):
    return {"arn": f"e{i}", "endpoints": list(endpoints)}


def lis(i, ports=(80,), proto="TCP", egs=None):
    return {"arn": f"l{i}", "proto": proto, "ports": list(ports), "egs": [eg(i)] if egs is None else list(egs)}


def acc(i=0, name="service-default-s", owner="service/default/s", host=HOST, enabled=True, listeners=None, extra_tags=()):
    return {"arn": f"a{i}", "name": name, "dns": f"a{i}.awsglobalaccelerator.com", "enabled": enabled,
            "tags": [("aws-global-accelerator-controller-managed", "true"), ("aws-global-accelerator-owner", owner),
                     ("aws-global-accelerator-target-hostname", host), ("aws-global-accelerator-cluster", "default")] + list(extra_tags),
            "listeners": [lis(i)] if listeners is None else list(listeners)}


def svc(ports=((80, "TCP"),), managed=True, lb_ingress=(HOST,), extra=None):
    ann = {"service.beta.kubernetes.io/aws-load-balancer-type": "nlb"}
    if managed:
        ann[ANN + "global-accelerator-managed"] = "true"
    ann.update(extra or {})
    return dict(kind="service", ns="default", name="s", spec_type="LoadBalancer", annotations=ann, ports=list(ports), lb_ingress=list(lb_ingress))


def ing(ports=(80,), extra=None, lb_ingress=(ALB_HOST,)):
    ann = {ANN + "global-accelerator-managed": "true"}
    ann.update(extra or {})
    return dict(kind="ingress", ns="default", name="i", ingress_class="alb", annotations=ann, ports=[(p, "TCP") for p in ports], lb_ingress=list(lb_ingress))


ALB = {"region": "us-west-2", "name": "k8s-default-i-0123456789", "dns": ALB_HOST, "arn": "arn:lb", "state": "active"}


def iacc(i=0, listeners=None):
    return acc(i, name="ingress-default-i", owner="ingress/default/i", host=ALB_HOST, listeners=listeners)


# name -> (object, lbs, accelerators, expected (status, detail, event), expected ops)
# an expected op is (opcode, a0, a1, a2) with rows of the flattened tables; lbIngress index (sub) is 0 throughout
V = {
    # ---- the load balancer gate
    "lb_not_found": (svc(), [], [acc()], (RETRY, D_LB_NOT_FOUND, 0), []),
    "lb_of_another_name": (svc(), [lb(name="f" * 32)], [acc()], (RETRY, D_LB_NOT_FOUND, 0), []),
    "lb_dns_name_differs": (svc(), [lb(dns="x" + HOST)], [acc()], (RETRY, D_LB_DNS, 0), []),
    "lb_not_active": (svc(), [lb(state="provisioning")], [], (REQ30, 0, 0), []),
    "lb_not_active_wins_over_a_pending_update": (svc(), [lb(state="active_impaired")], [acc(name="stale")], (REQ30, 0, 0), []),
    # ---- create
    "no_accelerator_creates_the_chain": (svc(), [lb()], [], (OK, 0, EV_CREATED), [(CREATE_CHAIN, 0, NONE, NONE)]),
    "accelerator_of_another_owner_does_not_count": (svc(), [lb()], [acc(owner="service/default/other")], (OK, 0, EV_CREATED), [(CREATE_CHAIN, 0, NONE, NONE)]),
    # ---- update, step by step
    "in_sync": (svc(), [lb()], [acc()], (OK, 0, 0), []),
    "accelerator_disabled": (svc(), [lb()], [acc(enabled=False)], (OK, 0, 0), [(UPDATE_ACCEL, 0, 0, NONE)]),
    "accelerator_name_differs": (svc(), [lb()], [acc(name="renamed")], (OK, 0, 0), [(UPDATE_ACCEL, 0, 0, NONE)]),
    "target_hostname_tag_differs": (svc(), [lb()], [acc(host="old-" + HOST)], (OK, 0, 0), [(UPDATE_ACCEL, 0, 0, NONE)]),
    "no_listener": (svc(), [lb()], [acc(listeners=[])], (OK, 0, 0), [(CREATE_LIS, 0, NONE, NONE), (CREATE_EG, 0, NONE, 0)]),
    "two_listeners_is_an_error": (svc(), [lb()], [acc(listeners=[lis(0), lis(1)])], (RETRY, D_MANY_LIS, 0), []),
    "ops_before_the_error_stay": (svc(), [lb()], [acc(name="renamed", listeners=[lis(0), lis(1)])], (RETRY, D_MANY_LIS, 0), [(UPDATE_ACCEL, 0, 0, NONE)]),
    "listener_protocol_differs": (svc(), [lb()], [acc(listeners=[lis(0, proto="UDP")])], (OK, 0, 0), [(UPDATE_LIS, 0, 0, NONE)]),
    "service_port_missing_on_the_listener": (svc(ports=((80, "TCP"), (443, "TCP"))), [lb()], [acc()], (OK, 0, 0), [(UPDATE_LIS, 0, 0, NONE)]),
    "listener_port_not_on_the_service": (svc(), [lb()], [acc(listeners=[lis(0, ports=(80, 443))])], (OK, 0, 0), [(UPDATE_LIS, 0, 0, NONE)]),
    "udp_service_last_protocol_wins": (svc(ports=((53, "TCP"), (53, "UDP"))), [lb()], [acc(listeners=[lis(0, ports=(53,), proto="UDP")])], (OK, 0, 0), []),
    "no_endpoint_group": (svc(), [lb()], [acc(listeners=[lis(0, egs=[])])], (OK, 0, 0), [(CREATE_EG, 0, 0, 0)]),
    "two_endpoint_groups_is_an_error": (svc(), [lb()], [acc(listeners=[lis(0, egs=[eg(0), eg(1)])])], (RETRY, D_MANY_EGS, 0), []),
    "listener_update_precedes_the_endpoint_group_error": (svc(), [lb()], [acc(listeners=[lis(0, proto="UDP", egs=[eg(0), eg(1)])])], (RETRY, D_MANY_EGS, 0), [(UPDATE_LIS, 0, 0, NONE)]),
    "endpoint_group_without_the_lb": (svc(), [lb()], [acc(listeners=[lis(0, egs=[eg(0, ["arn:other"])])])], (OK, 0, 0), [(UPDATE_EG, 0, 0, 0)]),
    "endpoint_group_empty": (svc(), [lb()], [acc(listeners=[lis(0, egs=[eg(0, [])])])], (OK, 0, 0), [(UPDATE_EG, 0, 0, 0)]),
    "lb_among_other_endpoints": (svc(), [lb()], [acc(listeners=[lis(0, egs=[eg(0, ["arn:other", "arn:lb"])])])], (OK, 0, 0), []),
    "everything_stale_in_statement_order": (svc(), [lb()], [acc(name="renamed", listeners=[lis(0, proto="UDP", egs=[eg(0, ["arn:other"])])])], (OK, 0, 0),
                                            [(UPDATE_ACCEL, 0, 0, NONE), (UPDATE_LIS, 0, 0, NONE), (UPDATE_EG, 0, 0, 0)]),
    "listener_update_then_endpoint_group_create": (svc(), [lb()], [acc(listeners=[lis(0, ports=(81,), egs=[])])], (OK, 0, 0), [(UPDATE_LIS, 0, 0, NONE), (CREATE_EG, 0, 0, 0)]),
    # ---- several accelerators of one owner: all of them, in list order
    "second_accelerator_is_updated_too": (svc(), [lb()], [acc(0), acc(1, name="renamed")], (OK, 0, 0), [(UPDATE_ACCEL, 1, 0, NONE)]),
    "both_accelerators_in_list_order": (svc(), [lb()], [acc(0, listeners=[lis(0, egs=[])]), acc(1, enabled=False)], (OK, 0, 0), [(CREATE_EG, 0, 0, 0), (UPDATE_ACCEL, 1, 0, NONE)]),
    "error_on_the_first_accelerator_ends_the_object": (svc(), [lb()], [acc(0, listeners=[lis(0), lis(1)]), acc(2, name="renamed")], (RETRY, D_MANY_LIS, 0), []),
    "another_owners_accelerator_in_between_is_skipped": (svc(), [lb()], [acc(0), acc(1, owner="service/default/other", name="x"), acc(2, name="renamed")], (OK, 0, 0), [(UPDATE_ACCEL, 2, 0, NONE)]),
    # ---- Ingress: listener protocol is always TCP, ports from the listen-ports annotation or the rules
    "ingress_in_sync": (ing(), [ALB], [iacc()], (OK, 0, 0), []),
    "ingress_udp_listener_is_always_wrong": (ing(), [ALB], [iacc(listeners=[lis(0, proto="UDP")])], (OK, 0, 0), [(UPDATE_LIS, 0, 0, NONE)]),
    "ingress_listen_ports_annotation_replaces_the_rules": (ing(ports=(80,), extra={"alb.ingress.kubernetes.io/listen-ports": '[{"HTTPS":443}]'}), [ALB], [iacc()], (OK, 0, 0), [(UPDATE_LIS, 0, 0, NONE)]),
    "ingress_listen_ports_annotation_in_sync": (ing(ports=(80,), extra={"alb.ingress.kubernetes.io/listen-ports": '[{"HTTP": 80}, {"HTTPS": 443}]'}), [ALB],
                                                [iacc(listeners=[lis(0, ports=(80, 443))])], (OK, 0, 0), []),
    "ingress_broken_listen_ports_means_no_ports": (ing(ports=(80,), extra={"alb.ingress.kubernetes.io/listen-ports": "not json"}), [ALB], [iacc()], (OK, 0, 0), [(UPDATE_LIS, 0, 0, NONE)]),
    # ---- no managed annotation: cleanup, every time
    "unmanaged_without_lb_ingress_returns_first": (svc(managed=False, lb_ingress=()), [lb()], [acc()], (SKIP_NO_LB, 0, 0), []),
    "unmanaged_nothing_to_delete": (svc(managed=False), [lb()], [], (OK, 0, EV_DELETED), []),
    "unmanaged_deletes_the_chain": (svc(managed=False), [lb()], [acc()], (OK, 0, EV_DELETED), [(DELETE_CHAIN, 0, 0, 0)]),
    "unmanaged_needs_no_load_balancer": (svc(managed=False), [], [acc()], (OK, 0, EV_DELETED), [(DELETE_CHAIN, 0, 0, 0)]),
    "cleanup_with_two_listeners_deletes_only_the_accelerator": (svc(managed=False), [lb()], [acc(listeners=[lis(0), lis(1)])], (OK, 0, EV_DELETED), [(DELETE_CHAIN, 0, NONE, NONE)]),
    "cleanup_without_listener": (svc(managed=False), [lb()], [acc(listeners=[])], (OK, 0, EV_DELETED), [(DELETE_CHAIN, 0, NONE, NONE)]),
    "cleanup_listener_without_endpoint_group": (svc(managed=False), [lb()], [acc(listeners=[lis(0, egs=[])])], (OK, 0, EV_DELETED), [(DELETE_CHAIN, 0, 0, NONE)]),
    "cleanup_with_two_endpoint_groups_keeps_them_out": (svc(managed=False), [lb()], [acc(listeners=[lis(0, egs=[eg(0), eg(1)])])], (OK, 0, EV_DELETED), [(DELETE_CHAIN, 0, 0, NONE)]),
    "cleanup_of_every_accelerator_of_the_owner": (svc(managed=False), [lb()], [acc(0), acc(1, owner="service/default/other"), acc(2)], (OK, 0, EV_DELETED),
                                                  [(DELETE_CHAIN, 0, 0, 0), (DELETE_CHAIN, 2, 2, 2)]),
    "managed_annotation_value_is_not_read": (svc(managed=False, extra={ANN + "global-accelerator-managed": "false"}), [lb()], [acc()], (OK, 0, 0), []),
}


def model(name):
    obj, lbs, accs, _, _ = V[name]
    return [obj], {"lbs": list(lbs), "accelerators": list(accs), "zones": []}


def observed(status_ga, section_begin, ops):
    w = int(status_ga[0])
    sb = [int(x) for x in section_begin]
    out = []
    for o in ops[sb[0]:sb[1]]:
        assert (int(o["head"]) >> 8) & 0xFF == 0 and int(o["obj"]) == 0 and int(o["sub"]) == 0
        out.append((int(o["head"]) & 0xFF, int(o["a0"]), int(o["a1"]), int(o["a2"])))
    return (w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF), out


def check(status_ga, section_begin, ops, name):
    status, got = observed(status_ga, section_begin, ops)
    assert status == V[name][3], (name, status)
    assert got == V[name][4], (name, got)


@pytest.mark.parametrize("name", sorted(V))
def test_oracle_matches_the_hand_derived_vectors(garecon, oracle, name):
    snap = garecon.pack(*model(name))
    for mode in (0, 1, 2):
        cs = oracle.diff(snap, "default", mode=mode)
        check(cs.status_ga, cs.section_begin, cs.ops, name)


@pytest.mark.parametrize("name", sorted(V))
def test_independent_python_restatement_matches(garecon, name):
    import importlib
    pyref = importlib.import_module("oracle.pyref")
    res = pyref.diff(*model(name), "default")
    check(res["status_ga"], res["section_begin"], [dict(zip(("head", "obj", "sub", "a0", "a1", "a2"), op)) for op in res["ops"]], name)


@pytest.mark.parametrize("name", sorted(V))
def test_device_logic_matches_the_hand_derived_vectors(garecon, name):
    import __graft_entry__ as ge
    snap = garecon.pack(*model(name))
    with garecon.Engine(cluster_name="default", lib=garecon.abi.load_library(ge.build_hostsim())) as e:
        e.load(snap)
        cs = e.diff()
        check(cs.status_ga, cs.section_begin, cs.ops, name)


@pytest.mark.gpu
def test_gpu_matches_the_hand_derived_vectors(garecon, engine):
    for name in sorted(V):
        engine.load(garecon.pack(*model(name)))
        cs = engine.diff()
        check(cs.status_ga, cs.section_begin, cs.ops, name)
