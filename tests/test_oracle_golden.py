"""Pins the CPU oracle against every known-answer vector the reference's own unit tests hold for the path
(SURVEY.md §8c).  The reference's tables live in tests/golden/reference_vectors.json (transcribed from the Go test files,
file:line per table); the cases further down that are NOT from the reference say so."""
import json
from pathlib import Path

import pytest

TCP, UDP = 0, 1
GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "reference_vectors.json").read_text())


def _cases(name):
    return [tuple(c) for c in GOLDEN[name]["cases"]]


# pkg/cloudprovider/aws/load_balancer_test.go:17-40  TestGetLBNameFromHostname
LB_HOSTNAMES = _cases("get_lb_name_from_hostname")


@pytest.mark.parametrize("hostname,name,region,code", LB_HOSTNAMES)
def test_get_lb_name_from_hostname(oracle, hostname, name, region, code):
    assert oracle.get_lb_name_from_hostname(hostname) == (code, name, region)


# pkg/cloudprovider/provider_test.go:14-19  TestDetectCloudProvider
@pytest.mark.parametrize("hostname,code", _cases("detect_cloud_provider"))
def test_detect_cloud_provider(oracle, hostname, code):
    assert oracle.detect_cloud_provider(hostname) == code


# hand-simulated vectors recorded in SURVEY.md §8.3 quirk 2 (derived from load_balancer.go:32-93)
@pytest.mark.parametrize("hostname,expect", [
    ("abc-.elb.us-east-1.amazonaws.com", (8, None, None)),
    ("nohyphen.elb.us-east-1.amazonaws.com", (8, None, None)),
    ("internal-x.us-east-1.elb.amazonaws.com", (6, None, None)),
    ("x-1.elb.amazonaws.com", (1, "x", "elb")),
    ("foo.example.amazonaws.com", (5, None, None)),
])
def test_hostname_quirks(oracle, hostname, expect):
    assert oracle.get_lb_name_from_hostname(hostname) == expect


def test_detect_cloud_provider_edges(oracle):
    assert oracle.detect_cloud_provider("example.com") == 1   # error -> `continue`
    assert oracle.detect_cloud_provider("localhost") == 2     # parts[len-2] panics
    assert oracle.detect_cloud_provider("") == 2
    assert oracle.detect_cloud_provider("amazonaws.com") == 0


# pkg/cloudprovider/aws/global_accelerator_test.go:22-146  TestListenerProtocolChange
PROTOCOL_CASES = _cases("listener_protocol_change")


@pytest.mark.parametrize("listener,protos,changed", PROTOCOL_CASES)
def test_listener_protocol_change(oracle, listener, protos, changed):
    assert (oracle.service_protocol(protos) != listener) == changed


# global_accelerator_test.go:164-335  TestListenerPortChanged
PORT_CASES = _cases("listener_port_changed")


@pytest.mark.parametrize("listener,svc,changed", PORT_CASES)
def test_listener_port_changed(oracle, listener, svc, changed):
    assert oracle.listener_port_changed(listener, svc) == changed


def test_listener_port_changed_quirks(oracle):
    # both empty -> unchanged; duplicates on one side mask a diff
    assert oracle.listener_port_changed([], []) is False
    assert oracle.listener_port_changed([80, 80], []) is False
    assert oracle.listener_port_changed([], [443]) is True


# global_accelerator_test.go:354-480  TestListenerForIngress — through the full oracle on a packed snapshot
def _ingress(garecon, annotations, ports):
    return garecon.pack([dict(kind="ingress", ns="default", name="Test ingress", ingress_class="alb", annotations=annotations, ports=ports)])


def test_listener_for_ingress(garecon, oracle):
    # "Only spec rules": rules-only -> [80]
    s = _ingress(garecon, {}, [80])
    cs = oracle.diff(s)
    assert not cs.derived[0] & garecon.abi.DV_PORTS_FROM_ANN and not cs.derived[0] & garecon.abi.DV_PROTO_UDP
    # "DefaultBackend is specified": packer order defaultBackend, then rules -> [8080, 80] (list is the input itself)
    s = _ingress(garecon, {}, [8080, 80])
    cs = oracle.diff(s)
    assert not cs.derived[0] & garecon.abi.DV_PORTS_FROM_ANN
    assert list(s.arrays["port_number"]) == [8080, 80]
    # "ListenPorts annotation is specified": annotation overrides the spec
    s = _ingress(garecon, {"alb.ingress.kubernetes.io/listen-ports": '[{"HTTP": 80}, {"HTTPS": 443}]'}, [8080, 80])
    cs = oracle.diff(s)
    assert cs.derived[0] & garecon.abi.DV_PORTS_FROM_ANN
    assert list(cs.dports) == [80, 443] and list(cs.dport_begin) == [0, 2]
    assert oracle.parse_listen_ports('[{"HTTP": 80}, {"HTTPS": 443}]') == [80, 443]


@pytest.mark.parametrize("val,ports", _cases("e2e_listen_ports"))
def test_e2e_listen_ports_fixture(oracle, val, ports):
    # local_e2e/pkg/fixtures/ingress.go:18 + local_e2e/e2e_test.go:192-205: exactly one port range 443-443
    assert oracle.parse_listen_ports(val) == ports


@pytest.mark.parametrize("annotations,backend_ports,ports,proto", _cases("listener_for_ingress"))
def test_listener_for_ingress_table(garecon, oracle, annotations, backend_ports, ports, proto):
    """The reference's three TestListenerForIngress cases through the full oracle on a packed one-Ingress snapshot."""
    s = _ingress(garecon, annotations, backend_ports)
    cs = oracle.diff(s)
    from_ann = bool(cs.derived[0] & garecon.abi.DV_PORTS_FROM_ANN)
    got = list(cs.dports) if from_ann else list(s.arrays["port_number"])
    assert got == ports and bool(cs.derived[0] & garecon.abi.DV_PROTO_UDP) == bool(proto)


# pkg/cloudprovider/aws/route53_test.go:19-85  TestFindARecord
A, CNAME = 1, 3
FIND_A = _cases("find_a_record")


@pytest.mark.parametrize("names,types,hostname,idx", FIND_A)
def test_find_a_record(oracle, names, types, hostname, idx):
    assert oracle.find_a_record(names, types, hostname) == idx


# route53_test.go:101-135  TestNeedRecordsUpdate
@pytest.mark.parametrize("has_alias,alias_dns,accel_dns,expected", _cases("need_records_update"))
def test_need_records_update(oracle, has_alias, alias_dns, accel_dns, expected):
    assert oracle.need_records_update(has_alias, alias_dns, accel_dns) is expected


# route53_test.go:150-175  TestParentDomain
@pytest.mark.parametrize("hostname,parent", _cases("parent_domain"))
def test_parent_domain(oracle, hostname, parent):
    assert oracle.parent_domain(hostname) == parent


# route53.go:18-20 (no reference test; literal format incl. the double quotes — SURVEY.md §8.3 quirk 8)
def test_route53_owner_value(oracle):
    assert oracle.route53_owner_value("default", "service", "ns", "n") == '"heritage=aws-global-accelerator-controller,cluster=default,service/ns/n"'


# Go encoding/json behaviours the annotation branch depends on (global_accelerator.go:526-542; quirk 4)
@pytest.mark.parametrize("val,ports", [
    ('[{"http": 80}]', [80]),                   # case-insensitive key match
    ('[{"HTTP": 80, "HTTPS": 443}]', [80, 443]),  # HTTP first, then HTTPS
    ('[{"HTTPS": 443, "HTTP": 80}]', [80, 443]),
    ('[{"HTTP": 0, "HTTPS": 443}]', [443]),     # zero is skipped
    ('[{"HTTP": 80, "extra": [1, {"a": null}]}]', [80]),  # unknown keys ignored
    ('[{"HTTP": 80}, null, {}]', [80]),
    ('null', []),
    ('[]', []),
    (' [ { "HTTP" : 80 } ] ', [80]),
    ('[{"HTTP": 4294967376}]', [80]),           # int64 -> int32 truncation
    ('[{"HTTP": 80, "HTTP": 81}]', [81]),       # duplicate key: last wins
    ('[{"HTTP": null}]', []),
    ('[{"\\u0048TTP": 8080}]', [8080]),         # escapes in keys are unquoted before matching
    ('[{"httpſ": 443}]', [443]),           # U+017F folds to S
    ('[{"HTTP": -5}]', [-5]),
])
def test_listen_ports_ok(oracle, val, ports):
    assert oracle.parse_listen_ports(val) == ports


@pytest.mark.parametrize("val", [
    '', '[', '[{"HTTP": 80}', '[{"HTTP": 80},]', '{"HTTP": 80}', '"x"', '80', 'true',
    '[{"HTTP": "80"}]', '[{"HTTP": 80.0}]', '[{"HTTP": 1e2}]', '[{"HTTP": true}]', '[{"HTTP": {}}]',
    '[{"HTTP": 9223372036854775808}]', '[80]', '[[{"HTTP":80}]]', '[{"HTTP": 080}]', '[{"HTTP": 80}] x',
    "[{'HTTP': 80}]", '[{"HTTP": 80, }]', '[{"HTTP": +80}]', '[{"HTTP": 80}]\x00', '[{"HT\tTP": 80}]',
    '[{"HTTP": 80, "x": "\\q"}]', '[{"HTTP": 80, "x": tru}]', '[{"HTTP": 80, "x": 1.}]', '[{"HTTP": 80, "x": -}]',
])
def test_listen_ports_error_means_no_ports(oracle, val):
    assert oracle.parse_listen_ports(val) is None


def test_json_depth_limit(oracle):
    # encoding/json scanner: maxNestingDepth = 10000
    ok = '[{"x": ' + '[' * 9998 + ']' * 9998 + ', "HTTP": 1}]'
    bad = '[{"x": ' + '[' * 9999 + ']' * 9999 + ', "HTTP": 1}]'
    assert oracle.parse_listen_ports(ok) == [1]
    assert oracle.parse_listen_ports(bad) is None
