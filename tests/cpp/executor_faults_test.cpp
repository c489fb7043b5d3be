// SURVEY.md §8 row f2, the failure half: host/executor.hpp on a MockCloud with injected AWS errors, through the C ABI
// (libgarecon.so on the GPU tier, the hostsim test build on the CPU tier).  Every expectation below is derived by hand from the
// reference: partial-create rollback (global_accelerator.go:139-147, :213-232), the Ingress path swallowing a listener error
// (:241-244), first error ends the object (service.go:107-110), events (service.go:82,116-118; route53/service.go:67,101-103),
// rate-limited requeue of failed keys (reconcile.go:75-77), and paginated ingestion in host/packer.hpp.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "aws-global-accelerator-controller_b200/host/executor.hpp"
#include "aws-global-accelerator-controller_b200/host/reconcile.hpp"

using namespace garecon;

static int g_fail = 0;
#define CHECK(cond)                                                       \
  do {                                                                    \
    if (!(cond)) {                                                        \
      fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond); \
      g_fail++;                                                           \
    }                                                                     \
  } while (0)

static const std::string ANN = kAnnPrefix;

static LoadBalancer nlb(int i) {
  char name[40], dns[120];
  snprintf(name, sizeof(name), "%032x", i);
  snprintf(dns, sizeof(dns), "%s-%016x.elb.us-west-2.amazonaws.com", name, i);
  LoadBalancer lb;
  lb.region = "us-west-2";
  lb.name = name;
  lb.dns = dns;
  lb.arn = "arn:aws:elasticloadbalancing:us-west-2:1:loadbalancer/net/" + std::string(name) + "/x";
  return lb;
}
static LoadBalancer alb(int i) {
  LoadBalancer lb;
  lb.region = "us-west-2";
  lb.name = "k8s-default-ing-" + std::to_string(i);
  lb.dns = lb.name + "-" + std::to_string(100000000 + i) + ".us-west-2.elb.amazonaws.com";
  lb.arn = "arn:aws:elasticloadbalancing:us-west-2:1:loadbalancer/app/" + lb.name + "/x";
  return lb;
}
static KObject svc(const std::string &name, const LoadBalancer &lb, const std::string &r53 = "") {
  KObject k;
  k.kind = GAR_KIND_SERVICE;
  k.ns = "default";
  k.name = name;
  k.annotations = {{"service.beta.kubernetes.io/aws-load-balancer-type", "nlb"}, {ANN + "global-accelerator-managed", "true"}};
  if (!r53.empty()) k.annotations.push_back({ANN + "route53-hostname", r53});
  k.lb_hostnames = {lb.dns};
  k.ports = {{80, "TCP"}, {443, "TCP"}};
  return k;
}
static KObject ing(const std::string &name, const LoadBalancer &lb) {
  KObject k;
  k.kind = GAR_KIND_INGRESS;
  k.ns = "default";
  k.name = name;
  k.has_ingress_class = true;
  k.ingress_class = "alb";
  k.annotations = {{ANN + "global-accelerator-managed", "true"}};
  k.lb_hostnames = {lb.dns};
  k.ports = {{80, ""}};
  return k;
}

struct Run {
  gar_engine *e;
  Packer p;
  gar_changeset cs{};
  explicit Run(gar_engine *eng) : e(eng) {}
  void diff(const std::vector<KObject> &objects, const CloudState &state) {
    gar_changeset_free(e, &cs);
    p.Reset();
    for (auto &k : objects) p.AddObject(k);
    p.AddCloud(state);
    p.Finish();
    if (gar_snapshot_load(e, p.objects(), p.actual()) != GAR_OK || gar_diff(e, &cs) != GAR_OK) {
      fprintf(stderr, "engine: %s\n", gar_last_error(e));
      exit(3);
    }
  }
  ~Run() { gar_changeset_free(e, &cs); }
};
static std::string joined(const std::vector<std::string> &v) {
  std::string s;
  for (auto &x : v) s += (s.empty() ? "" : ",") + x;
  return s;
}
static size_t count_events(const ExecReport &r, const char *reason) {
  size_t n = 0;
  for (auto &e : r.events) n += e.reason == reason;
  return n;
}

struct Queue : RateLimitingQueue {
  std::deque<std::string> pending;
  std::vector<std::string> log;
  bool Get(std::string *key, bool *shutdown) override {
    *shutdown = false;
    if (pending.empty()) return false;
    *key = pending.front();
    pending.pop_front();
    return true;
  }
  void Done(const std::string &) override {}
  void Forget(const std::string &k) override { log.push_back("forget:" + k); }
  void AddRateLimited(const std::string &k) override { log.push_back("ratelimited:" + k); }
  void AddAfter(const std::string &k, std::chrono::nanoseconds) override { log.push_back("after:" + k); }
};

int main() {
  const char *cluster = "default";
  gar_config ec{GAR_ABI_VERSION, 0, cluster, 0};
  gar_engine *e = nullptr;
  if (gar_engine_create(&ec, &e) != GAR_OK) {
    fprintf(stderr, "engine: %s\n", gar_last_error(nullptr));
    return 2;
  }
  {  // 1. Service: createListener fails after the accelerator exists -> rollback, error, no event; the next batch creates again
    std::vector<KObject> objs = {svc("a", nlb(1)), svc("b", nlb(2))};
    CloudState st;
    st.lbs = {nlb(1), nlb(2)};
    Run r(e);
    r.diff(objs, st);
    CHECK(r.cs.n_ops == 2);
    MockCloud cloud(&st);
    cloud.FailNth("CreateListener", 1);
    ExecReport rep = ExecuteChangeSet(r.cs, objs, st, cluster, cloud);
    CHECK(joined(cloud.calls) == "CreateAccelerator,CreateListener!,DeleteAccelerator,CreateAccelerator,CreateListener,CreateEndpointGroup");
    CHECK(rep.Failed(GAR_CTRL_GA, 0) && !rep.Failed(GAR_CTRL_GA, 1) && rep.rollbacks == 1 && rep.executed == 1);
    CHECK(count_events(rep, "GlobalAcceleratorCreated") == 1 && rep.events[0].obj == 1);
    cloud.Commit();
    CHECK(st.accelerators.size() == 1 && st.accelerators[0].name == "service-default-b");
    r.diff(objs, st);  // object a is created by the next batch
    CHECK(r.cs.n_ops == 1 && (r.cs.ops[0].head & 0xFF) == GAR_OP_GA_CREATE_CHAIN && r.cs.ops[0].obj == 0);
  }
  {  // 2. Service: createEndpointGroup fails -> the listener and the accelerator made so far are cleaned up
    std::vector<KObject> objs = {svc("a", nlb(1))};
    CloudState st;
    st.lbs = {nlb(1)};
    Run r(e);
    r.diff(objs, st);
    MockCloud cloud(&st);
    cloud.FailNth("CreateEndpointGroup", 1);
    ExecReport rep = ExecuteChangeSet(r.cs, objs, st, cluster, cloud);
    CHECK(joined(cloud.calls) == "CreateAccelerator,CreateListener,CreateEndpointGroup!,DeleteListener,DeleteAccelerator");
    CHECK(rep.Failed(GAR_CTRL_GA, 0) && count_events(rep, "GlobalAcceleratorCreated") == 0);
    CHECK(count_events(rep, "Route53RecordDeleted") == 1);  // no route53 annotation: CleanupRecordSet ran (nothing to delete) and the reference records the event every time (route53/service.go:54-69)
    cloud.Commit();
    CHECK(st.accelerators.empty());
  }
  {  // 3. Ingress: the same listener error is SWALLOWED (createGlobalAcceleratorForIngress returns arn, nil): success, event,
     //    a listener-less accelerator stays; the next batch adds listener + endpoint group through the update path
    std::vector<KObject> objs = {ing("i", alb(1))};
    CloudState st;
    st.lbs = {alb(1)};
    Run r(e);
    r.diff(objs, st);
    MockCloud cloud(&st);
    cloud.FailNth("CreateListener", 1);
    ExecReport rep = ExecuteChangeSet(r.cs, objs, st, cluster, cloud);
    CHECK(joined(cloud.calls) == "CreateAccelerator,CreateListener!");
    CHECK(!rep.Failed(GAR_CTRL_GA, 0) && rep.rollbacks == 0 && count_events(rep, "GlobalAcceleratorCreated") == 1);
    cloud.Commit();
    CHECK(st.accelerators.size() == 1 && st.accelerators[0].listeners.empty());
    r.diff(objs, st);
    CHECK(r.cs.n_ops == 2 && (r.cs.ops[0].head & 0xFF) == GAR_OP_GA_CREATE_LISTENER && (r.cs.ops[1].head & 0xFF) == GAR_OP_GA_CREATE_EG);
    MockCloud cloud2(&st);
    rep = ExecuteChangeSet(r.cs, objs, st, cluster, cloud2);
    cloud2.Commit();
    CHECK(rep.executed == 2 && st.accelerators[0].listeners.size() == 1 && st.accelerators[0].listeners[0].endpoint_groups.size() == 1);
    r.diff(objs, st);
    CHECK(r.cs.n_ops == 0);
  }
  {  // 4. the first error ends the object's op stream; other objects and the other controller are unaffected
    std::vector<KObject> objs = {svc("a", nlb(1), "a.example.com"), svc("b", nlb(2))};
    CloudState st;
    st.lbs = {nlb(1), nlb(2)};
    st.zones = {HostedZone{"Z1", "example.com.", {}}};
    for (int i = 0; i < 2; i++) {  // both accelerators exist but are disabled and listen on the wrong ports: UPDATE_ACCEL + UPDATE_LISTENER each
      Accelerator a;
      a.arn = "acc" + std::to_string(i);
      a.name = std::string("service-default-") + (i ? "b" : "a");
      a.dns = a.arn + ".awsglobalaccelerator.com";
      a.enabled = false;
      a.tags = {{kTagManaged, "true"}, {kTagOwner, std::string("service/default/") + (i ? "b" : "a")}, {kTagTargetHostname, st.lbs[i].dns}, {kTagCluster, cluster}};
      Listener l;
      l.arn = a.arn + "/l";
      l.from_ports = {80};
      l.endpoint_groups = {EndpointGroup{l.arn + "/e", {st.lbs[i].arn}}};
      a.listeners = {l};
      st.accelerators.push_back(a);
    }
    Run r(e);
    r.diff(objs, st);
    CHECK(r.cs.section_begin[1] == 4 && r.cs.section_begin[3] - r.cs.section_begin[2] == 1);  // 2 + 2 GA ops, one R53_CREATE
    MockCloud cloud(&st);
    cloud.FailNth("UpdateAccelerator", 1);
    ExecReport ga = ExecuteChangeSet(r.cs, objs, st, cluster, cloud, kExecGA);
    CHECK(joined(cloud.calls) == "UpdateAccelerator!,UpdateAccelerator,UpdateListener");  // object a: its listener update is skipped
    CHECK(ga.Failed(GAR_CTRL_GA, 0) && !ga.Failed(GAR_CTRL_GA, 1) && ga.skipped == 1 && ga.executed == 2);
    ExecReport r53 = ExecuteChangeSet(r.cs, objs, st, cluster, cloud, kExecR53);  // the route53 worker executes its own half only
    CHECK(r53.executed == 1 && !r53.Failed(GAR_CTRL_R53, 0) && count_events(r53, "Route53RecourdCreated") == 1);
    CHECK(cloud.calls.size() == 5 && cloud.calls[3] == "CreateMetadataRecordSet" && cloud.calls[4] == "CreateRecordSet");
    cloud.Commit();
    CHECK(!st.accelerators[0].enabled && st.accelerators[1].enabled && st.accelerators[0].listeners[0].from_ports.size() == 1);
    // the failed key is requeued rate-limited, the healthy one forgotten (reconcile.go:75-77,86-88)
    r.diff(objs, st);
    Queue q;
    q.pending = {"default/a", "default/b", "default/gone"};
    MockCloud cloud2(&st);
    cloud2.FailNth("UpdateAccelerator", 1);
    BatchStats bs;
    int rc = ProcessBatch(
        e, q, Controller::GlobalAccelerator, GAR_KIND_SERVICE, [&](const std::string &k) -> int64_t { return k == "default/a" ? 0 : k == "default/b" ? 1 : -1; },
        [&](const gar_changeset &cs) {
          ExecReport rep = ExecuteChangeSet(cs, objs, st, cluster, cloud2, kExecGA);
          return OpFailures{rep.failed, rep.failed_owners};
        },
        &bs);
    CHECK(rc == GAR_OK && joined(q.log) == "ratelimited:default/a,forget:default/b,forget:default/gone");
    for (auto &c : cloud2.calls) CHECK(c.find("Record") == std::string::npos);  // the GA worker did not touch route53
  }
  {  // 5. route53: the TXT record is created, the A record fails -> error, no Created event (ensureRoute53 returns false, err)
    std::vector<KObject> objs = {svc("a", nlb(1), "a.example.com,b.example.com")};
    CloudState st;
    st.lbs = {nlb(1)};
    st.zones = {HostedZone{"Z1", "example.com.", {}}};
    Accelerator a;
    a.arn = "acc";
    a.name = "service-default-a";
    a.dns = "acc.awsglobalaccelerator.com";
    a.tags = {{kTagManaged, "true"}, {kTagOwner, "service/default/a"}, {kTagTargetHostname, st.lbs[0].dns}, {kTagCluster, cluster}};
    Listener l;
    l.from_ports = {80, 443};
    l.endpoint_groups = {EndpointGroup{"e", {st.lbs[0].arn}}};
    a.listeners = {l};
    st.accelerators = {a};
    Run r(e);
    r.diff(objs, st);
    CHECK(r.cs.n_ops == 2);
    MockCloud cloud(&st);
    cloud.FailNth("CreateRecordSet", 2);  // the second hostname's A record
    ExecReport rep = ExecuteChangeSet(r.cs, objs, st, cluster, cloud);
    CHECK(joined(cloud.calls) == "CreateMetadataRecordSet,CreateRecordSet,CreateMetadataRecordSet,CreateRecordSet!");
    CHECK(rep.Failed(GAR_CTRL_R53, 0) && count_events(rep, "Route53RecourdCreated") == 0);
  }
  {  // 6. cleanup of a key that left the cache fails -> that key is requeued, other orphans are still cleaned
    std::vector<KObject> objs = {svc("keep", nlb(1))};
    CloudState st;
    st.lbs = {nlb(1)};
    for (const char *owner : {"service/default/gone1", "service/default/gone2", "service/default/keep"}) {
      Accelerator a;
      a.arn = std::string("acc-") + owner;
      a.name = "x";
      a.dns = "x.awsglobalaccelerator.com";
      a.tags = {{kTagManaged, "true"}, {kTagOwner, owner}, {kTagTargetHostname, st.lbs[0].dns}, {kTagCluster, cluster}};
      Listener l;
      l.from_ports = {80, 443};
      l.endpoint_groups = {EndpointGroup{"e", {st.lbs[0].arn}}};
      a.listeners = {l};
      st.accelerators.push_back(a);
    }
    st.accelerators[2].name = "service-default-keep";
    Run r(e);
    r.diff(objs, st);
    CHECK(r.cs.section_begin[2] - r.cs.section_begin[1] == 2);  // two orphan delete chains
    MockCloud cloud(&st);
    cloud.FailNth("DeleteListener", 1);
    ExecReport rep = ExecuteChangeSet(r.cs, objs, st, cluster, cloud);
    CHECK(joined(cloud.calls) == "DeleteEndpointGroup,DeleteListener!,DeleteEndpointGroup,DeleteListener,DeleteAccelerator");
    CHECK(rep.failed_owners.count("service/default/gone1") == 1 && rep.failed_owners.size() == 1);
    Queue q;
    q.pending = {"default/gone1", "default/gone2"};
    MockCloud cloud2(&st);
    cloud2.FailNth("DeleteListener", 1);
    ProcessBatch(
        e, q, Controller::GlobalAccelerator, GAR_KIND_SERVICE, [&](const std::string &) -> int64_t { return -1; },
        [&](const gar_changeset &cs) {
          ExecReport rp = ExecuteChangeSet(cs, objs, st, cluster, cloud2, kExecGA);
          return OpFailures{rp.failed, rp.failed_owners};
        });
    CHECK(joined(q.log) == "ratelimited:default/gone1,forget:default/gone2");
  }
  {  // 7. GAR_PENDING arguments: two lbIngress, nothing exists -> ONE accelerator, re-tagged and re-pointed to the second LB
    KObject k = svc("a", nlb(1));
    k.lb_hostnames.push_back(nlb(2).dns);
    std::vector<KObject> objs = {k};
    CloudState st;
    st.lbs = {nlb(1), nlb(2)};
    Run r(e);
    r.diff(objs, st);
    CHECK(r.cs.n_ops == 3 && r.cs.ops[1].a0 == GAR_PENDING && r.cs.ops[2].a1 == GAR_PENDING);
    MockCloud cloud(&st);
    ExecReport rep = ExecuteChangeSet(r.cs, objs, st, cluster, cloud);
    cloud.Commit();
    CHECK(rep.executed == 3 && st.accelerators.size() == 1);
    CHECK(st.accelerators[0].listeners[0].endpoint_groups[0].endpoint_ids == std::vector<std::string>{st.lbs[1].arn});
    bool thost2 = false;
    for (auto &t : st.accelerators[0].tags) thost2 = thost2 || (t.first == kTagTargetHostname && t.second == st.lbs[1].dns);
    CHECK(thost2);
  }
  {  // 8. paginated ingestion (host/packer.hpp): pages in list order produce exactly the tables AddCloud produces
    CloudState st;
    for (int i = 0; i < 7; i++) st.lbs.push_back(nlb(i));
    for (int i = 0; i < 5; i++) {
      Accelerator a;
      a.name = "n" + std::to_string(i);
      a.dns = "d" + std::to_string(i);
      a.tags = {{kTagOwner, "service/default/o" + std::to_string(i)}};
      Listener l;
      l.from_ports = {80 + i};
      l.endpoint_groups = {EndpointGroup{"e", {"arn" + std::to_string(i)}}};
      a.listeners = {l};
      st.accelerators.push_back(a);
    }
    for (int z = 0; z < 4; z++) {
      HostedZone hz{"Z" + std::to_string(z), "z" + std::to_string(z) + ".example.com.", {}};
      for (int rr = 0; rr < (z == 2 ? 0 : 5); rr++) {
        RecordSet rs;
        rs.name = "r" + std::to_string(rr) + "." + hz.name;
        rs.type = GAR_RR_TXT;
        rs.values = {"v" + std::to_string(rr), "w"};
        hz.records.push_back(rs);
      }
      st.zones.push_back(hz);
    }
    Packer whole, paged;
    whole.AddCloud(st);
    whole.Finish();
    paged.AddLoadBalancerPage({st.lbs.begin(), st.lbs.begin() + 4});
    paged.AddLoadBalancerPage({st.lbs.begin() + 4, st.lbs.end()});
    CHECK(paged.AddAcceleratorPage({st.accelerators.begin(), st.accelerators.begin() + 2}) == 0);
    CHECK(paged.AddAcceleratorPage({st.accelerators.begin() + 2, st.accelerators.end()}) == 2);
    paged.AddHostedZonePage({st.zones.begin(), st.zones.begin() + 3});
    paged.AddHostedZonePage({st.zones.begin() + 3, st.zones.end()});
    CHECK(paged.AddRecordSetPage(0, {st.zones[0].records.begin(), st.zones[0].records.begin() + 3}));
    CHECK(paged.AddRecordSetPage(0, {st.zones[0].records.begin() + 3, st.zones[0].records.end()}));
    CHECK(paged.AddRecordSetPage(1, st.zones[1].records));
    CHECK(paged.AddRecordSetPage(3, st.zones[3].records));   // zone 2 has no record sets
    CHECK(!paged.AddRecordSetPage(1, st.zones[1].records));  // going back would break the zone-major row order
    paged.Finish();
    const gar_actual *a = whole.actual(), *b = paged.actual();
    CHECK(a->n_lbs == b->n_lbs && a->n_accels == b->n_accels && a->n_zones == b->n_zones && a->n_records == b->n_records && a->n_values == b->n_values);
    CHECK(a->slab_len == b->slab_len && !memcmp(a->slab, b->slab, a->slab_len));
    CHECK(!memcmp(a->zone_rec_begin, b->zone_rec_begin, 4 * (a->n_zones + 1)) && !memcmp(a->rec_val_begin, b->rec_val_begin, 4 * (a->n_records + 1)));
    CHECK(!memcmp(a->rec_name, b->rec_name, 8 * a->n_records) && !memcmp(a->val_value, b->val_value, 8 * a->n_values) && !memcmp(a->lb_dns, b->lb_dns, 8 * a->n_lbs));
    CHECK(!memcmp(a->acc_lis_begin, b->acc_lis_begin, 4 * (a->n_accels + 1)) && !memcmp(a->ep_id, b->ep_id, 8 * a->n_endpoints));
  }
  {  // 9. convergence THROUGH failures: every batch loses some calls, the account still ends where the fault-free run ends
    auto build = [&](std::vector<KObject> &objs, CloudState &st) {
      st.zones = {HostedZone{"Z1", "example.com.", {}}};
      for (int i = 0; i < 40; i++) {
        bool ingress = i % 3 == 0;
        LoadBalancer lb = ingress ? alb(i) : nlb(i);
        st.lbs.push_back(lb);
        KObject k = ingress ? ing("o" + std::to_string(i), lb) : svc("o" + std::to_string(i), lb, i % 2 ? "h" + std::to_string(i) + ".example.com" : "");
        objs.push_back(k);
      }
    };
    auto summary = [&](const CloudState &st) {
      size_t lis = 0, egs = 0, recs = st.zones[0].records.size();
      for (auto &a : st.accelerators) {
        lis += a.listeners.size();
        for (auto &l : a.listeners) egs += l.endpoint_groups.size();
      }
      return std::to_string(st.accelerators.size()) + "/" + std::to_string(lis) + "/" + std::to_string(egs) + "/" + std::to_string(recs);
    };
    std::string want;
    for (int faulty = 0; faulty < 2; faulty++) {
      std::vector<KObject> objs;
      CloudState st;
      build(objs, st);
      Run r(e);
      int rounds = 0;
      for (; rounds < 12; rounds++) {
        r.diff(objs, st);
        if (r.cs.n_ops == 0) break;
        MockCloud cloud(&st);
        if (faulty && rounds < 6) {
          cloud.FailNth("CreateListener", 2 + rounds);
          cloud.FailNth("CreateEndpointGroup", 3);
          cloud.FailNth("CreateRecordSet", 1 + rounds);
          cloud.FailNth("CreateAccelerator", 5);
        }
        ExecuteChangeSet(r.cs, objs, st, cluster, cloud);
        cloud.Commit();
      }
      CHECK(rounds < 12);
      if (!faulty) want = summary(st);
      else {
        // a TXT record whose A record failed stays behind and the retry writes both again: compare accelerators/listeners/EGs
        // exactly and require at least the fault-free number of records
        std::string got = summary(st);
        CHECK(got.substr(0, got.rfind('/')) == want.substr(0, want.rfind('/')));
        CHECK(std::stoul(got.substr(got.rfind('/') + 1)) >= std::stoul(want.substr(want.rfind('/') + 1)));
      }
    }
  }
  gar_engine_destroy(e);
  printf("{\"failed_checks\": %d}\n", g_fail);
  return g_fail ? 1 : 0;
}
