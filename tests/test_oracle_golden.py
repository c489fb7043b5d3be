"""Pins the CPU oracle against every known-answer vector the reference's own unit tests hold for the path
(SURVEY.md §8c).  Each table below is the reference's table, cited file:line."""
import pytest

TCP, UDP = 0, 1


# pkg/cloudprovider/aws/load_balancer_test.go:17-40  TestGetLBNameFromHostname
LB_HOSTNAMES = [
    ("aa5849cde256f49faa7487bb433155b7-3f43353a6cb6f633.elb.ap-northeast-1.amazonaws.com", "aa5849cde256f49faa7487bb433155b7", "ap-northeast-1", 2),
    ("test-b6cdc5fbd1d6fa43.elb.ap-northeast-1.amazonaws.com", "test", "ap-northeast-1", 2),
    ("k8s-default-h3poteto-f1f41628db-201899272.ap-northeast-1.elb.amazonaws.com", "k8s-default-h3poteto-f1f41628db", "ap-northeast-1", 1),
    ("internal-k8s-default-h3poteto-35ca57562f-777774719.ap-northeast-1.elb.amazonaws.com", "k8s-default-h3poteto-35ca57562f", "ap-northeast-1", 0),
]


@pytest.mark.parametrize("hostname,name,region,code", LB_HOSTNAMES)
def test_get_lb_name_from_hostname(oracle, hostname, name, region, code):
    assert oracle.get_lb_name_from_hostname(hostname) == (code, name, region)


# pkg/cloudprovider/provider_test.go:14-19  TestDetectCloudProvider
def test_detect_cloud_provider(oracle):
    assert oracle.detect_cloud_provider("aa5849cde256f49faa7487bb433155b7-3f43353a6cb6f633.elb.ap-northeast-1.amazonaws.com") == 0


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
PROTOCOL_CASES = [
    (UDP, ["UDP"], False),
    (TCP, ["TCP", "TCP"], False),
    (TCP, ["UDP", "TCP"], False),
    (TCP, ["UDP"], True),
    (TCP, ["UDP", "UDP"], True),
    (TCP, ["TCP", "UDP"], True),
]


@pytest.mark.parametrize("listener,protos,changed", PROTOCOL_CASES)
def test_listener_protocol_change(oracle, listener, protos, changed):
    assert (oracle.service_protocol(protos) != listener) == changed


# global_accelerator_test.go:164-335  TestListenerPortChanged
PORT_CASES = [
    ([80], [80], False),
    ([80, 443, 8080], [443, 8080, 80], False),
    ([80], [443], True),
    ([80, 8080], [443, 8080], True),
    ([80, 8080], [443, 8080, 8081], True),
    ([80, 443, 8080], [443], True),
]


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


def test_e2e_listen_ports_fixture(oracle):
    # local_e2e/pkg/fixtures/ingress.go:18 + local_e2e/e2e_test.go:192-205: exactly one port range 443-443
    assert oracle.parse_listen_ports('[{"HTTPS":443}]') == [443]


# pkg/cloudprovider/aws/route53_test.go:19-85  TestFindARecord
A, CNAME = 1, 3
FIND_A = [
    (["foo.example.com.", "bar.example.com."], [CNAME, CNAME], "foo.example.com", -1),
    (["foo.example.com.", "bar.example.com."], [A, A], "baz.example.com", -1),
    (["foo.example.com.", "bar.example.com."], [A, A], "bar.example.com", 1),
    (["\\052.example.com.", "bar.example.com."], [A, A], "*.example.com", 0),
]


@pytest.mark.parametrize("names,types,hostname,idx", FIND_A)
def test_find_a_record(oracle, names, types, hostname, idx):
    assert oracle.find_a_record(names, types, hostname) == idx


# route53_test.go:101-135  TestNeedRecordsUpdate
def test_need_records_update(oracle):
    assert oracle.need_records_update(False, "", "") is True
    assert oracle.need_records_update(True, "foo.example.com.", "bar.example.com") is True
    assert oracle.need_records_update(True, "foo.example.com.", "foo.example.com") is False


# route53_test.go:150-175  TestParentDomain
@pytest.mark.parametrize("hostname,parent", [
    ("h3poteto-test.example.com", "example.com"),
    ("h3poteto-test.foo.example.com", "foo.example.com"),
    ("example.com", "com"),
    ("com", ""),
    (".", ""),
])
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
