"""Hand-derived vectors for the Route53 half of the path, which the reference pins only through findARecord / needRecordsUpdate /
parentDomain unit tables (SURVEY.md §8c "parity unpinned": a9 ensure, a10 cleanup).  Every expectation below is derived from the
Go text, not from any of the restatements:

  ensureRoute53 (route53.go:56-130)
    * ListGlobalAcceleratorByHostname must return exactly one accelerator, else Result{Requeue, 1m} and nothing is touched (:68-77);
    * per hostname of strings.Split(annotation, ",") — no trimming: GetHostedZone, FindOwneredARecordSets, findARecord, then
      create (TXT owner record, then A alias) / nothing / UPSERT of the A alias;
  GetHostedZone (:335-360): walk hostname, parentDomain(hostname), ... ; a zone matches when its Name == target + "." exactly; the
      empty string ends the walk with an error ("Could not find hosted zone") -> AddRateLimited;
  FindOwneredARecordSets (:216-238): names of the record sets holding a value == the owner value, then every set with one of those
      names AND an AliasTarget (any type);
  findARecord (:362-369): first of those with Type A and replaceWildcards(Name) == hostname + "." (first "\\052" becomes "*");
  needRecordsUpdate (:375-383): alias DNSName != accelerator.DnsName + ".";
  Route53OwnerValue (:18-20): "\"heritage=aws-global-accelerator-controller,cluster=<cluster>,<resource>/<ns>/<name>\"";
  route53/service.go:54-69: no route53-hostname annotation -> CleanupRecordSet + Route53RecordDeleted event, every time;
  CleanupRecordSet (:132-165): per zone, the owned alias sets first, then the metadata sets — once PER MATCHING VALUE.

One Service `default/s` behind an NLB, one accelerator tagged for its hostname; the inputs of a vector are the route53 annotation,
the accelerators and the zones; the observable is the R53 status word and the R53 ops."""
import pytest

ANN = "aws-global-accelerator-controller.h3poteto.dev/"
HOST = "0123456789abcdef0123456789abcdef-0123456789abcdef.elb.us-west-2.amazonaws.com"
LB = {"region": "us-west-2", "name": "0123456789abcdef0123456789abcdef", "dns": HOST, "arn": "arn:lb", "state": "active"}
ACC_DNS = "a0123456789abcdef.awsglobalaccelerator.com"
OWNER = '"heritage=aws-global-accelerator-controller,cluster=default,service/default/s"'
OTHER = '"heritage=aws-global-accelerator-controller,cluster=default,service/default/other"'

OK, REQ60, RETRY = 1, 4, 5
D_NO_ZONE = 9
EV_CREATED, EV_DELETED = 1, 2
CREATE, UPSERT, DELETE = 8, 9, 10


def acc(i=0, host=HOST, cluster="default", managed="true"):
    return {"arn": f"a{i}", "name": f"service-default-s{i}", "dns": ACC_DNS if i == 0 else f"b{i}.awsglobalaccelerator.com", "enabled": True,
            "tags": [("aws-global-accelerator-controller-managed", managed), ("aws-global-accelerator-owner", "service/default/s"),
                     ("aws-global-accelerator-target-hostname", host), ("aws-global-accelerator-cluster", cluster)],
            "listeners": [{"arn": f"l{i}", "proto": "TCP", "ports": [80], "egs": [{"arn": f"e{i}", "endpoints": ["arn:lb"]}]}]}


def txt(name, *values):
    return {"name": name, "type": "TXT", "values": list(values)}


def alias(name, target=ACC_DNS + ".", rtype="A"):
    return {"name": name, "type": rtype, "alias": target}


def zone(name, *records):
    return {"id": "/hostedzone/Z" + name, "name": name, "records": list(records)}


IN_SYNC = [txt("app.example.com.", OWNER), alias("app.example.com.")]

# name -> (route53 annotation or None, accelerators, zones, lb_ingress, expected status (code, detail, event), expected ops)
# an expected op is (opcode, k = hostname index, zone row, record row within the flattened record table or None)
V = {
    # ---- the accelerator lookup gate
    "no_accelerator": ("app.example.com", [], [zone("example.com.")], [HOST], (REQ60, 11, 0), []),
    "two_accelerators": ("app.example.com", [acc(0), acc(1)], [zone("example.com.")], [HOST], (REQ60, 10, 0), []),
    "accelerator_of_another_cluster_is_not_listed": ("app.example.com", [acc(0, cluster="prod")], [zone("example.com.")], [HOST], (REQ60, 11, 0), []),
    "accelerator_for_another_hostname_is_not_listed": ("app.example.com", [acc(0, host="x" + HOST)], [zone("example.com.")], [HOST], (REQ60, 11, 0), []),
    # ---- GetHostedZone
    "zone_missing_is_an_error": ("app.example.com", [acc()], [zone("example.org.")], [HOST], (RETRY, D_NO_ZONE, 0), []),
    "zone_found_by_walking_to_the_parent": ("app.example.com", [acc()], [zone("example.org."), zone("example.com.")], [HOST], (OK, 0, EV_CREATED), [(CREATE, 0, 1, None)]),
    "zone_walk_reaches_the_tld": ("app.example.com", [acc()], [zone("com.")], [HOST], (OK, 0, EV_CREATED), [(CREATE, 0, 0, None)]),
    "exact_zone_wins_over_the_parent": ("app.example.com", [acc()], [zone("example.com."), zone("app.example.com.")], [HOST], (OK, 0, EV_CREATED), [(CREATE, 0, 1, None)]),
    "zone_name_without_trailing_dot_never_matches": ("app.example.com", [acc()], [zone("example.com")], [HOST], (RETRY, D_NO_ZONE, 0), []),
    "zone_name_case_matters": ("app.example.com", [acc()], [zone("Example.com.")], [HOST], (RETRY, D_NO_ZONE, 0), []),
    # ---- findARecord / needRecordsUpdate over FindOwneredARecordSets
    "in_sync": ("app.example.com", [acc()], [zone("example.com.", *IN_SYNC)], [HOST], (OK, 0, 0), []),
    "alias_points_elsewhere": ("app.example.com", [acc()], [zone("example.com.", txt("app.example.com.", OWNER), alias("app.example.com.", "stale.awsglobalaccelerator.com."))], [HOST],
                               (OK, 0, 0), [(UPSERT, 0, 0, 1)]),
    "alias_without_the_trailing_dot_is_stale": ("app.example.com", [acc()], [zone("example.com.", txt("app.example.com.", OWNER), alias("app.example.com.", ACC_DNS))], [HOST],
                                                (OK, 0, 0), [(UPSERT, 0, 0, 1)]),
    "plain_a_record_is_not_an_owned_alias": ("app.example.com", [acc()], [zone("example.com.", txt("app.example.com.", OWNER), {"name": "app.example.com.", "type": "A", "values": ["1.2.3.4"]})], [HOST],
                                             (OK, 0, EV_CREATED), [(CREATE, 0, 0, None)]),
    "owner_record_of_another_object": ("app.example.com", [acc()], [zone("example.com.", txt("app.example.com.", OTHER), alias("app.example.com."))], [HOST],
                                       (OK, 0, EV_CREATED), [(CREATE, 0, 0, None)]),
    "owner_value_among_other_values": ("app.example.com", [acc()], [zone("example.com.", txt("app.example.com.", '"v=spf1 -all"', OWNER), alias("app.example.com."))], [HOST], (OK, 0, 0), []),
    "aaaa_alias_is_not_an_a_record": ("app.example.com", [acc()], [zone("example.com.", txt("app.example.com.", OWNER), alias("app.example.com.", rtype="AAAA"))], [HOST],
                                      (OK, 0, EV_CREATED), [(CREATE, 0, 0, None)]),
    "first_matching_a_alias_decides": ("app.example.com", [acc()], [zone("example.com.", txt("app.example.com.", OWNER), alias("app.example.com.", "stale.awsglobalaccelerator.com."), alias("app.example.com."))], [HOST],
                                       (OK, 0, 0), [(UPSERT, 0, 0, 1)]),
    "record_name_case_matters": ("app.example.com", [acc()], [zone("example.com.", txt("App.example.com.", OWNER), alias("App.example.com."))], [HOST], (OK, 0, EV_CREATED), [(CREATE, 0, 0, None)]),
    "owner_txt_under_another_name_does_not_own_this_alias": ("app.example.com", [acc()], [zone("example.com.", txt("other.example.com.", OWNER), alias("app.example.com."))], [HOST],
                                                             (OK, 0, EV_CREATED), [(CREATE, 0, 0, None)]),
    "records_of_another_zone_are_not_seen": ("app.example.com", [acc()], [zone("example.com."), zone("example.org.", *IN_SYNC)], [HOST], (OK, 0, EV_CREATED), [(CREATE, 0, 0, None)]),
    # ---- wildcards: Route53 stores "*" as "\052"
    "wildcard_in_sync": ("*.example.com", [acc()], [zone("example.com.", txt("\\052.example.com.", OWNER), alias("\\052.example.com."))], [HOST], (OK, 0, 0), []),
    "wildcard_literal_star_record_also_matches": ("*.example.com", [acc()], [zone("example.com.", txt("*.example.com.", OWNER), alias("*.example.com."))], [HOST], (OK, 0, 0), []),
    "only_the_first_escape_is_replaced": ("*.*.example.com", [acc()], [zone("example.com.", txt("\\052.\\052.example.com.", OWNER), alias("\\052.\\052.example.com."))], [HOST],
                                          (OK, 0, EV_CREATED), [(CREATE, 0, 0, None)]),
    # ---- the hostname list: split on "," only
    "two_hostnames_second_missing": ("app.example.com,api.example.com", [acc()], [zone("example.com.", *IN_SYNC)], [HOST], (OK, 0, EV_CREATED), [(CREATE, 1, 0, None)]),
    "two_hostnames_both_missing_in_order": ("app.example.com,api.example.org", [acc()], [zone("example.org."), zone("example.com.")], [HOST], (OK, 0, EV_CREATED),
                                            [(CREATE, 0, 1, None), (CREATE, 1, 0, None)]),
    "no_trimming_of_the_pieces": ("app.example.com, api.example.com", [acc()], [zone("example.com.", *IN_SYNC, txt("api.example.com.", OWNER), alias("api.example.com."))], [HOST],
                                  (OK, 0, EV_CREATED), [(CREATE, 1, 0, None)]),  # " api.example.com" (leading blank) is another name: zone by walking, no such record
    "first_hostname_without_zone_stops_the_object": ("app.example.net,app.example.com", [acc()], [zone("example.com.")], [HOST], (RETRY, D_NO_ZONE, 0), []),
    "ops_before_the_error_stay": ("app.example.com,app.example.net", [acc()], [zone("example.com.")], [HOST], (RETRY, D_NO_ZONE, 0), [(CREATE, 0, 0, None)]),
    "empty_annotation_is_one_empty_hostname": ("", [acc()], [zone("example.com.")], [HOST], (RETRY, D_NO_ZONE, 0), []),
    # ---- the lbIngress loop
    "no_lb_ingress": ("app.example.com", [acc()], [zone("example.com.")], [], (OK, 0, 0), []),
    "not_an_aws_hostname_is_skipped": ("app.example.com", [acc()], [zone("example.com.")], ["lb.example.net"], (OK, 0, 0), []),
    # ---- no annotation: cleanup on every reconcile, Route53RecordDeleted every time
    "no_annotation_nothing_owned": (None, [acc()], [zone("example.com.", txt("app.example.com.", OTHER), alias("app.example.com."))], [HOST], (OK, 0, EV_DELETED), []),
    "no_annotation_cleanup_alias_then_metadata": (None, [acc()], [zone("example.com.", *IN_SYNC)], [HOST], (OK, 0, EV_DELETED), [(DELETE, 0, 0, 1), (DELETE, 1, 0, 0)]),
    "cleanup_visits_every_zone_in_list_order": (None, [acc()], [zone("example.com.", *IN_SYNC), zone("example.org.", txt("x.example.org.", OWNER), alias("x.example.org.", "anything."))], [HOST],
                                                (OK, 0, EV_DELETED), [(DELETE, 0, 0, 1), (DELETE, 1, 0, 0), (DELETE, 0, 1, 3), (DELETE, 1, 1, 2)]),
    "metadata_set_deleted_once_per_matching_value": (None, [acc()], [zone("example.com.", txt("app.example.com.", OWNER, OWNER))], [HOST], (OK, 0, EV_DELETED), [(DELETE, 1, 0, 0), (DELETE, 1, 0, 0)]),
    "cleanup_deletes_owned_aliases_of_any_type": (None, [acc()], [zone("example.com.", txt("app.example.com.", OWNER), alias("app.example.com.", rtype="AAAA"), alias("app.example.com."))], [HOST],
                                                  (OK, 0, EV_DELETED), [(DELETE, 0, 0, 1), (DELETE, 0, 0, 2), (DELETE, 1, 0, 0)]),
    "cleanup_needs_no_accelerator": (None, [], [zone("example.com.", *IN_SYNC)], [HOST], (OK, 0, EV_DELETED), [(DELETE, 0, 0, 1), (DELETE, 1, 0, 0)]),
}


def model(name):
    ann_v, accs, zones, ingress, _, _ = V[name]
    ann = {"service.beta.kubernetes.io/aws-load-balancer-type": "nlb"}
    if ann_v is not None:
        ann[ANN + "route53-hostname"] = ann_v
    obj = dict(kind="service", ns="default", name="s", spec_type="LoadBalancer", annotations=ann, ports=[(80, "TCP")], lb_ingress=list(ingress))
    return [obj], {"lbs": [LB], "accelerators": list(accs), "zones": list(zones)}


def observed(cs):
    w = int(cs.status_r53[0])
    sb = [int(x) for x in cs.section_begin]
    # section 3 (orphan sweep) holds the cleanup of `service/default/other`, whose records some vectors plant and which is not in
    # the cache; it is not what these vectors are about (tests/test_gpu_parity.py::test_orphans_only and the randomized models cover it)
    ops = []
    for o in cs.ops[sb[2]:sb[3]]:
        code = int(o["head"]) & 0xFF
        assert (int(o["head"]) >> 8) & 0xFF == 1 and int(o["obj"]) == 0
        if code == DELETE:
            ops.append((code, int(o["sub"]), int(o["a0"]), int(o["a1"])))           # phase, zone, record
        else:
            assert int(o["sub"]) >> 20 == 0 and int(o["a1"]) == 0                      # lbIngress 0, the one accelerator
            ops.append((code, int(o["sub"]) & 0xFFFFF, int(o["a0"]), int(o["a2"]) if code == UPSERT else None))
    return (w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF), ops


def check(cs, name):
    want_status, want_ops = V[name][4], V[name][5]
    status, ops = observed(cs)
    assert status == want_status, (name, status)
    assert ops == want_ops, (name, ops)


@pytest.mark.parametrize("name", sorted(V))
def test_oracle_matches_the_hand_derived_vectors(garecon, oracle, name):
    snap = garecon.pack(*model(name))
    for mode in (0, 1, 2):
        check(oracle.diff(snap, "default", mode=mode), name)


@pytest.mark.parametrize("name", sorted(V))
def test_independent_python_restatement_matches(garecon, name):
    import importlib
    pyref = importlib.import_module("oracle.pyref")
    objects, actual = model(name)
    res = pyref.diff(objects, actual, "default")

    class CS:
        status_r53 = res["status_r53"]
        section_begin = res["section_begin"]
        ops = [dict(zip(("head", "obj", "sub", "a0", "a1", "a2"), op)) for op in res["ops"]]
    check(CS, name)


@pytest.mark.parametrize("name", sorted(V))
def test_device_logic_matches_the_hand_derived_vectors(garecon, name):
    import __graft_entry__ as ge
    snap = garecon.pack(*model(name))
    with garecon.Engine(cluster_name="default", lib=garecon.abi.load_library(ge.build_hostsim())) as e:
        e.load(snap)
        check(e.diff(), name)


@pytest.mark.gpu
def test_gpu_matches_the_hand_derived_vectors(garecon, engine):
    for name in sorted(V):
        engine.load(garecon.pack(*model(name)))
        check(engine.diff(), name)
