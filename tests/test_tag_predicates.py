"""Hand-derived vectors for the two tag predicates the reference has no unit test for (SURVEY.md §8c "parity unpinned"):
tagsContainsAllValues (global_accelerator.go:559-570) inside ListGlobalAcceleratorByResource (:87-110) and acceleratorChanged
(:412-437) with acceleratorTags (:35-51).  Every expectation is derived from the Go text:

  * `actual[*t.Key] = *t.Value` in list order  -> a later duplicate of a tag key wins;
  * `actual[k] != v` with a missing key        -> reads "" (a target value "" matches a missing tag);
  * acceleratorTags: strings.Split(annotation, ","), then strings.Split(piece, "="), pieces with len != 2 are dropped
    ("a=b=c", "bad", "" all vanish; "=v" and "k=" are kept with an empty key / value);
  * targetTags[k] = v in annotation order after the three system tags -> a user tag replaces a system tag, a later piece an earlier.

One Service, one load balancer, one accelerator whose tag list is the test's input; the observable is whether the change set
holds GA_UPDATE_ACCEL (acceleratorChanged), GA_CREATE_CHAIN (not listed for the owner) or nothing."""
import pytest

ANN = "aws-global-accelerator-controller.h3poteto.dev/"
HOST = "0123456789abcdef0123456789abcdef-0123456789abcdef.elb.us-west-2.amazonaws.com"
LB = {"region": "us-west-2", "name": "0123456789abcdef0123456789abcdef", "dns": HOST, "arn": "arn:lb", "state": "active"}
M, O, H, C = ("aws-global-accelerator-controller-managed", "aws-global-accelerator-owner", "aws-global-accelerator-target-hostname", "aws-global-accelerator-cluster")
SYS = [(M, "true"), (O, "service/default/s"), (H, HOST), (C, "default")]
NOTHING, UPDATE, CREATE = "nothing", "GA_UPDATE_ACCEL", "GA_CREATE_CHAIN"

# name -> (tags annotation or None, the accelerator's tag list, expected outcome)
VECTORS = {
    "in_sync": (None, SYS, NOTHING),
    # tagsContainsAllValues: later duplicate wins
    "duplicate_owner_last_matches": (None, [(O, "service/default/other")] + SYS, NOTHING),
    "duplicate_owner_last_differs": (None, SYS + [(O, "service/default/other")], CREATE),       # not listed for this owner -> create
    "duplicate_hostname_last_differs": (None, SYS + [(H, "stale")], UPDATE),
    "duplicate_managed_last_false": (None, SYS + [(M, "false")], CREATE),
    # missing tag reads as ""
    "cluster_tag_missing": (None, SYS[:3], CREATE),                                              # "" != "default" -> not listed
    "hostname_tag_missing": (None, [SYS[0], SYS[1], SYS[3]], UPDATE),
    "user_tag_with_empty_value_matches_missing_tag": ("k=", SYS, NOTHING),                        # actual["k"] == "" == target
    "user_tag_with_empty_key": ("=v", SYS + [("", "v")], NOTHING),
    "user_tag_with_empty_key_missing": ("=v", SYS, UPDATE),
    # acceleratorTags: only pieces with exactly one '='
    "three_parts_dropped": ("a=b=c", SYS, NOTHING),
    "no_equals_dropped": ("bad", SYS, NOTHING),
    "empty_annotation": ("", SYS, NOTHING),                                                       # Split("", ",") = [""] -> len 1 -> dropped
    "good_after_bad_piece": ("a=b=c,k=v", SYS, UPDATE),
    "good_after_bad_piece_present": ("a=b=c,k=v", SYS + [("k", "v")], NOTHING),
    "later_piece_replaces_earlier": ("k=v,k=w", SYS + [("k", "w")], NOTHING),
    "later_piece_replaces_earlier_stale": ("k=v,k=w", SYS + [("k", "v")], UPDATE),
    # a user tag replaces a system tag in targetTags
    "user_tag_overrides_hostname": (H + "=pinned", [SYS[0], SYS[1], (H, "pinned"), SYS[3]], NOTHING),
    "user_tag_overrides_hostname_actual_is_lb": (H + "=pinned", SYS, UPDATE),
    "user_tag_overrides_owner": (O + "=evil", SYS, UPDATE),                                       # listed (tags say s) but target owner is "evil"
    "user_tag_overrides_managed_same_value": (M + "=true", SYS, NOTHING),
    # the cluster tag is NOT part of acceleratorChanged's target (:420-424), but a user tag can put it there
    "user_tag_names_cluster_other_value": (C + "=prod", SYS, UPDATE),
    # case matters (Go string compare)
    "managed_value_case": (None, [(M, "True")] + SYS[1:], CREATE),
}


def model(tags_ann, acc_tags):
    ann = {ANN + "global-accelerator-managed": "true", "service.beta.kubernetes.io/aws-load-balancer-type": "nlb"}
    if tags_ann is not None:
        ann[ANN + "global-accelerator-tags"] = tags_ann
    obj = dict(kind="service", ns="default", name="s", spec_type="LoadBalancer", annotations=ann, ports=[(80, "TCP")], lb_ingress=[HOST])
    acc = {"arn": "a", "name": "service-default-s", "dns": "a.awsglobalaccelerator.com", "enabled": True, "tags": list(acc_tags),
           "listeners": [{"arn": "l", "proto": "TCP", "ports": [80], "egs": [{"arn": "e", "endpoints": ["arn:lb"]}]}]}
    return [obj], {"lbs": [LB], "accelerators": [acc], "zones": []}


def outcome(cs):
    sb = [int(x) for x in cs.section_begin]
    codes = [int(o["head"]) & 0xFF for o in cs.ops[sb[0]:sb[1]]]
    assert codes in ([], [1], [2]), codes
    return {(): NOTHING, (1,): CREATE, (2,): UPDATE}[tuple(codes)]


@pytest.mark.parametrize("name", sorted(VECTORS))
def test_oracle_matches_the_hand_derived_vectors(garecon, oracle, name):
    tags_ann, acc_tags, want = VECTORS[name]
    snap = garecon.pack(*model(tags_ann, acc_tags))
    for mode in (0, 1, 2):
        assert outcome(oracle.diff(snap, "default", mode=mode)) == want, (name, mode)


@pytest.mark.parametrize("name", sorted(VECTORS))
def test_device_logic_matches_the_hand_derived_vectors(garecon, name):
    import __graft_entry__ as ge
    tags_ann, acc_tags, want = VECTORS[name]
    snap = garecon.pack(*model(tags_ann, acc_tags))
    with garecon.Engine(cluster_name="default", lib=garecon.abi.load_library(ge.build_hostsim())) as e:
        e.load(snap)
        assert outcome(e.diff()) == want, name


@pytest.mark.gpu
def test_gpu_matches_the_hand_derived_vectors(garecon, engine):
    for name, (tags_ann, acc_tags, want) in sorted(VECTORS.items()):
        engine.load(garecon.pack(*model(tags_ann, acc_tags)))
        assert outcome(engine.diff()) == want, name
