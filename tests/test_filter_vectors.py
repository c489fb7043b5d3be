"""Hand-derived vectors for the row classifier: the controllers' event filters and the per-object derived
desired state, which the reference never unit-tests.  Derived from the Go text:

  wasLoadBalancerService (globalaccelerator/service.go:18-26, route53/service.go:19-27): spec.type == LoadBalancer AND (the
      service.beta.kubernetes.io/aws-load-balancer-type annotation is PRESENT — any value — OR spec.loadBalancerClass != nil);
  wasALBIngress (globalaccelerator/ingress.go:19-27): *spec.ingressClassName == "alb" (exactly) OR the kubernetes.io/ingress.class
      annotation is PRESENT — any value;
  hasManagedAnnotation / hasHostnameAnnotation (controller.go:250-253, route53/controller.go:243-246): presence only;
  the Route53 controller filters Services by wasLoadBalancerService (route53/controller.go:87-110) but takes EVERY Ingress
      (:129-148) — no class check;
  client-ip-preservation: annotation == "true" exactly (global_accelerator.go:225); ip-address-type: "ipv4" | "IPV4" (:679-681);
  listenerForService: the LAST port whose lower-cased protocol is tcp / udp decides (:503-515);
  listenerForIngress: the listen-ports annotation, when PRESENT, replaces the rules' ports (:526-542).

(CPU tiers only: oracle modes 0/1/2, pyref, and the device row logic through hostsim; the GPU tier classifies the same rows in
every parity test and compares the `derived` words bit for bit.)

Objects have no lbIngress, so an eligible object gets as far as the "no ingress LoadBalancer" return of the GA controller
(service.go:59-62 -> GAR_ST_SKIP_NO_LB) and the empty lbIngress loop / the cleanup branch of the Route53 controller."""
import pytest

ANN = "aws-global-accelerator-controller.h3poteto.dev/"
LBTYPE = "service.beta.kubernetes.io/aws-load-balancer-type"
ICLASS = "kubernetes.io/ingress.class"
UDP, IPPRES, IPV4, PORTS_ANN, GA_EL, GA_MAN, R53_EL, R53_ANN = (1 << i for i in range(8))
IGNORED, OK, SKIP_NO_LB = 0, 1, 2
EV_DELETED = 2


def svc(spec_type="LoadBalancer", ann=None, lb_class=False, ports=((80, "TCP"),)):
    return dict(kind="service", ns="default", name="s", spec_type=spec_type, lb_class=lb_class, annotations=dict(ann or {}), ports=list(ports), lb_ingress=[])


def ing(cls=None, ann=None, ports=(80,)):
    d = dict(kind="ingress", ns="default", name="i", annotations=dict(ann or {}), ports=[(p, "TCP") for p in ports], lb_ingress=[])
    if cls is not None:
        d["ingress_class"] = cls
    return d


# name -> (object, derived word, GA status code, (R53 status code, R53 event))
V = {
    # ---- Service: wasLoadBalancerService
    "lb_service_with_type_annotation": (svc(ann={LBTYPE: "nlb"}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "type_annotation_value_is_not_read": (svc(ann={LBTYPE: ""}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "lb_class_instead_of_the_annotation": (svc(lb_class=True), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "lb_service_with_neither": (svc(), 0, IGNORED, (IGNORED, 0)),
    "cluster_ip_with_the_annotation": (svc("ClusterIP", ann={LBTYPE: "nlb"}), 0, IGNORED, (IGNORED, 0)),
    "node_port_with_lb_class": (svc("NodePort", lb_class=True), 0, IGNORED, (IGNORED, 0)),
    "external_name": (svc("ExternalName", ann={LBTYPE: "nlb"}), 0, IGNORED, (IGNORED, 0)),
    "ineligible_service_annotations_do_not_matter": (svc(ann={ANN + "global-accelerator-managed": "true", ANN + "route53-hostname": "a.example.com"}), GA_MAN | R53_ANN, IGNORED, (IGNORED, 0)),
    # ---- presence-only annotations
    "managed_annotation_present": (svc(ann={LBTYPE: "nlb", ANN + "global-accelerator-managed": "true"}), GA_EL | GA_MAN | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "managed_annotation_empty_value_counts": (svc(ann={LBTYPE: "nlb", ANN + "global-accelerator-managed": ""}), GA_EL | GA_MAN | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "managed_annotation_false_counts_too": (svc(ann={LBTYPE: "nlb", ANN + "global-accelerator-managed": "false"}), GA_EL | GA_MAN | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "hostname_annotation_present": (svc(ann={LBTYPE: "nlb", ANN + "route53-hostname": "a.example.com"}), GA_EL | R53_EL | R53_ANN, SKIP_NO_LB, (OK, 0)),
    "annotation_key_prefix_is_not_enough": (svc(ann={LBTYPE: "nlb", ANN + "global-accelerator-managed-x": "true", ANN + "route53-hostnames": "a"}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "annotation_key_case_matters": (svc(ann={LBTYPE: "nlb", ANN + "Global-Accelerator-Managed": "true"}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    # ---- derived desired state
    "client_ip_preservation_true": (svc(ann={LBTYPE: "nlb", ANN + "client-ip-preservation": "true"}), GA_EL | R53_EL | IPPRES, SKIP_NO_LB, (OK, EV_DELETED)),
    "client_ip_preservation_is_case_sensitive": (svc(ann={LBTYPE: "nlb", ANN + "client-ip-preservation": "True"}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "ip_address_type_ipv4": (svc(ann={LBTYPE: "nlb", ANN + "ip-address-type": "ipv4"}), GA_EL | R53_EL | IPV4, SKIP_NO_LB, (OK, EV_DELETED)),
    "ip_address_type_upper": (svc(ann={LBTYPE: "nlb", ANN + "ip-address-type": "IPV4"}), GA_EL | R53_EL | IPV4, SKIP_NO_LB, (OK, EV_DELETED)),
    "ip_address_type_mixed_case_is_unknown": (svc(ann={LBTYPE: "nlb", ANN + "ip-address-type": "Ipv4"}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "ip_address_type_dualstack": (svc(ann={LBTYPE: "nlb", ANN + "ip-address-type": "dualstack"}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "udp_port_last": (svc(ann={LBTYPE: "nlb"}, ports=((53, "TCP"), (53, "UDP"))), GA_EL | R53_EL | UDP, SKIP_NO_LB, (OK, EV_DELETED)),
    "tcp_port_last": (svc(ann={LBTYPE: "nlb"}, ports=((53, "UDP"), (53, "TCP"))), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "protocol_is_lower_cased_first": (svc(ann={LBTYPE: "nlb"}, ports=((53, "Udp"),)), GA_EL | R53_EL | UDP, SKIP_NO_LB, (OK, EV_DELETED)),
    "sctp_does_not_change_the_protocol": (svc(ann={LBTYPE: "nlb"}, ports=((53, "UDP"), (54, "SCTP"))), GA_EL | R53_EL | UDP, SKIP_NO_LB, (OK, EV_DELETED)),
    "no_ports_is_tcp": (svc(ann={LBTYPE: "nlb"}, ports=()), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    # ---- Ingress: wasALBIngress for the GA controller, nothing for the Route53 controller
    "ingress_class_name_alb": (ing("alb"), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "ingress_class_name_is_case_sensitive": (ing("ALB"), R53_EL, IGNORED, (OK, EV_DELETED)),
    "ingress_class_name_nginx": (ing("nginx"), R53_EL, IGNORED, (OK, EV_DELETED)),
    "ingress_class_name_empty_string": (ing(""), R53_EL, IGNORED, (OK, EV_DELETED)),
    "ingress_without_any_class": (ing(), R53_EL, IGNORED, (OK, EV_DELETED)),
    "ingress_class_annotation_any_value": (ing(ann={ICLASS: "nginx"}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "ingress_class_annotation_beats_another_class_name": (ing("nginx", ann={ICLASS: ""}), GA_EL | R53_EL, SKIP_NO_LB, (OK, EV_DELETED)),
    "route53_takes_every_ingress": (ing("nginx", ann={ANN + "route53-hostname": "a.example.com"}), R53_EL | R53_ANN, IGNORED, (OK, 0)),
    "ingress_listen_ports_annotation_present": (ing("alb", ann={"alb.ingress.kubernetes.io/listen-ports": '[{"HTTP": 80}]'}), GA_EL | R53_EL | PORTS_ANN, SKIP_NO_LB, (OK, EV_DELETED)),
    "ingress_listen_ports_annotation_broken_is_still_present": (ing("alb", ann={"alb.ingress.kubernetes.io/listen-ports": "x"}), GA_EL | R53_EL | PORTS_ANN, SKIP_NO_LB, (OK, EV_DELETED)),
    "ingress_managed_and_ip_preserve": (ing("alb", ann={ANN + "global-accelerator-managed": "yes", ANN + "client-ip-preservation": "true"}), GA_EL | GA_MAN | R53_EL | IPPRES, SKIP_NO_LB, (OK, EV_DELETED)),
}


def check(cs, name):
    _, derived, st_ga, (st_r53, ev_r53) = V[name]
    assert int(cs.derived[0]) == derived, (name, bin(int(cs.derived[0])))
    assert int(cs.status_ga[0]) & 0xFF == st_ga, (name, hex(int(cs.status_ga[0])))
    w = int(cs.status_r53[0])
    assert (w & 0xFF, (w >> 16) & 0xFF) == (st_r53, ev_r53), (name, hex(w))
    assert len(cs.ops) == 0


def model(name):
    return [V[name][0]], {"lbs": [], "accelerators": [], "zones": []}


@pytest.mark.parametrize("name", sorted(V))
def test_oracle_matches_the_hand_derived_vectors(garecon, oracle, name):
    snap = garecon.pack(*model(name))
    for mode in (0, 1, 2):
        check(oracle.diff(snap, "default", mode=mode), name)


@pytest.mark.parametrize("name", sorted(V))
def test_independent_python_restatement_matches(garecon, name):
    import importlib
    pyref = importlib.import_module("oracle.pyref")
    res = pyref.diff(*model(name), "default")

    class CS:
        derived, status_ga, status_r53, ops = res["derived"], res["status_ga"], res["status_r53"], res["ops"]
    check(CS, name)


@pytest.mark.parametrize("name", sorted(V))
def test_device_logic_matches_the_hand_derived_vectors(garecon, name):
    import __graft_entry__ as ge
    snap = garecon.pack(*model(name))
    with garecon.Engine(cluster_name="default", lib=garecon.abi.load_library(ge.build_hostsim()), allow_empty_cache=True) as e:
        e.load(snap)
        check(e.diff(), name)
