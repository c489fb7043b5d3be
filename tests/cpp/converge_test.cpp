// Batch worker end to end in C++ (SURVEY.md §8 rows f1 + f2): a model of the informer cache and of AWS ->
// host/packer.hpp -> the C ABI (libgarecon.so on the GPU tier, the hostsim test build on the CPU tier) -> the change set,
// checked against the oracle bit for bit -> host/executor.hpp on the in-memory MockCloud -> next round.
// Expectation (what the reference's controllers do over successive reconciles): an empty AWS account converges to one
// accelerator per managed object and one TXT + A pair per annotated hostname in two rounds (accelerators, then records, as
// route53's ensure needs the accelerator to exist: route53.go:68-77 requeues until it does); injected drift and deleted
// objects are repaired / cleaned up; a converged state yields no ops.  Prints one JSON line per phase.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "aws-global-accelerator-controller_b200/host/executor.hpp"

extern "C" {
int orc_diff(const gar_objects *o, const gar_actual *a, const char *cluster, int mode, int threads, gar_changeset **out);
void orc_free(gar_changeset *cs);
}

using namespace garecon;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return rng_state;
}
static std::string hex(uint64_t v, int n) {
  char b[32];
  snprintf(b, sizeof(b), "%0*llx", n, (unsigned long long)v);
  return std::string(b).substr(0, (size_t)n);
}

struct Round {
  uint64_t n_ops = 0;
  size_t not_ok_ga = 0, not_ok_r53 = 0;
  bool oracle_equal = false;
  double pack_ms = 0, diff_ms = 0;
  size_t executed = 0, events = 0;
};

static Round run_round(gar_engine *e, const std::vector<KObject> &objects, CloudState &state, const char *cluster) {
  Round r;
  auto t0 = std::chrono::steady_clock::now();
  Packer p;
  for (auto &k : objects) p.AddObject(k);
  p.AddCloud(state);
  p.Finish();
  auto t1 = std::chrono::steady_clock::now();
  r.pack_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  if (gar_snapshot_load(e, p.objects(), p.actual()) != GAR_OK) {
    fprintf(stderr, "load: %s\n", gar_last_error(e));
    exit(3);
  }
  gar_changeset cs{};
  if (gar_diff(e, &cs) != GAR_OK) {
    fprintf(stderr, "diff: %s\n", gar_last_error(e));
    exit(4);
  }
  r.diff_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
  gar_changeset *want = nullptr;
  orc_diff(p.objects(), p.actual(), cluster, 1, 4, &want);
  r.oracle_equal = want->n_ops == cs.n_ops && want->n_objects == cs.n_objects &&
                   (cs.n_ops == 0 || !memcmp(want->ops, cs.ops, sizeof(gar_op) * cs.n_ops)) &&
                   (cs.n_objects == 0 || (!memcmp(want->status_ga, cs.status_ga, 4 * (size_t)cs.n_objects) && !memcmp(want->status_r53, cs.status_r53, 4 * (size_t)cs.n_objects)));
  orc_free(want);
  r.n_ops = cs.n_ops;
  for (uint32_t i = 0; i < cs.n_objects; i++) {
    uint32_t g = GAR_STATUS_CODE(cs.status_ga[i]), d = GAR_STATUS_CODE(cs.status_r53[i]);
    r.not_ok_ga += !(g == GAR_ST_OK || g == GAR_ST_IGNORED || g == GAR_ST_SKIP_NO_LB);
    r.not_ok_r53 += !(d == GAR_ST_OK || d == GAR_ST_IGNORED || d == GAR_ST_SKIP_NO_LB);
  }
  MockCloud cloud(&state);
  // the two controllers' workers each execute their own half of the change set
  ExecReport ga = ExecuteChangeSet(cs, objects, state, cluster, cloud, kExecGA);
  ExecReport r53 = ExecuteChangeSet(cs, objects, state, cluster, cloud, kExecR53);
  r.executed = ga.executed + r53.executed;
  r.events = ga.events.size() + r53.events.size();
  if (!ga.failed.empty() || !r53.failed.empty() || ga.executed + r53.executed != cs.n_ops) {
    fprintf(stderr, "executor: %zu + %zu of %llu ops executed without fault injection\n", ga.executed, r53.executed, (unsigned long long)cs.n_ops);
    exit(5);
  }
  cloud.Commit();
  gar_changeset_free(e, &cs);
  return r;
}

static void report(const char *phase, int round, const Round &r, const CloudState &s) {
  size_t nrec = 0;
  for (auto &z : s.zones) nrec += z.records.size();
  printf("{\"phase\": \"%s\", \"round\": %d, \"n_ops\": %llu, \"not_ok_ga\": %zu, \"not_ok_r53\": %zu, \"oracle_equal\": %s, \"accelerators\": %zu, \"records\": %zu, "
         "\"pack_ms\": %.3f, \"load_diff_ms\": %.3f}\n",
         phase, round, (unsigned long long)r.n_ops, r.not_ok_ga, r.not_ok_r53, r.oracle_equal ? "true" : "false", s.accelerators.size(), nrec, r.pack_ms, r.diff_ms);
}

int main(int argc, char **argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 200;
  const char *cluster = "default";
  const std::string ann = kAnnPrefix;
  std::vector<KObject> objects;
  CloudState state;
  const char *regions[] = {"us-west-2", "eu-west-1", "ap-northeast-1"};
  for (uint32_t z = 0; z < 4; z++) state.zones.push_back(HostedZone{"Z" + std::to_string(z), "z" + std::to_string(z) + ".example.com.", {}});
  for (uint32_t i = 0; i < n; i++) {
    KObject k;
    bool ingress = i % 3 == 1;
    k.kind = ingress ? GAR_KIND_INGRESS : GAR_KIND_SERVICE;
    k.ns = "team-" + std::to_string(i % 7);
    k.name = (ingress ? "ing-" : "svc-") + std::to_string(i);
    std::string region = regions[i % 3];
    LoadBalancer lb;
    lb.region = region;
    if (ingress) {
      k.has_ingress_class = true;
      k.ingress_class = "alb";
      lb.name = "k8s-" + k.ns + "-" + std::to_string(i);
      lb.dns = lb.name + "-" + std::to_string(100000000 + rnd() % 900000000) + "." + region + ".elb.amazonaws.com";
      lb.arn = "arn:aws:elasticloadbalancing:" + region + ":1:loadbalancer/app/" + lb.name + "/" + hex(rnd(), 16);
      k.ports = {{80, ""}};
      if (i % 2) k.annotations.push_back({"alb.ingress.kubernetes.io/listen-ports", "[{\"HTTP\": 80}, {\"HTTPS\": 443}]"});
    } else {
      k.annotations.push_back({"service.beta.kubernetes.io/aws-load-balancer-type", "nlb"});
      lb.name = hex(rnd(), 16) + hex(rnd(), 16);
      lb.dns = lb.name + "-" + hex(rnd(), 16) + ".elb." + region + ".amazonaws.com";
      lb.arn = "arn:aws:elasticloadbalancing:" + region + ":1:loadbalancer/net/" + lb.name + "/" + hex(rnd(), 16);
      k.ports = {{80, "TCP"}, {443, i % 11 == 0 ? "UDP" : "TCP"}};
    }
    if (i % 13 != 5) k.annotations.push_back({ann + "global-accelerator-managed", "yes"});
    if (i % 4 != 3) {
      std::string hosts = "h" + std::to_string(i) + ".z" + std::to_string(i % 4) + ".example.com";
      if (i % 5 == 0) hosts += ",*.w" + std::to_string(i) + ".z" + std::to_string((i + 1) % 4) + ".example.com";
      k.annotations.push_back({ann + "route53-hostname", hosts});
    }
    if (i % 9 == 0) k.annotations.push_back({ann + "global-accelerator-tags", "Environment=prod,Team=" + k.ns});
    if (i % 17 == 0) k.annotations.push_back({ann + "global-accelerator-name", "custom-" + std::to_string(i)});
    if (i % 19 == 7) lb.state = GAR_LB_PROVISIONING;
    k.lb_hostnames = {lb.dns};
    state.lbs.push_back(lb);
    objects.push_back(k);
  }
  gar_config ec{GAR_ABI_VERSION, 0, cluster, 0};
  gar_engine *e = nullptr;
  if (gar_engine_create(&ec, &e) != GAR_OK) {
    fprintf(stderr, "engine: %s\n", gar_last_error(nullptr));
    return 2;
  }
  // phase 1: empty account -> converge
  for (int round = 0; round < 5; round++) {
    Round r = run_round(e, objects, state, cluster);
    report("create", round, r, state);
    if (r.n_ops == 0) break;
  }
  // phase 2: the provisioning LBs become active; drift on some accelerators
  for (auto &lb : state.lbs) lb.state = GAR_LB_ACTIVE;
  for (size_t a = 0; a < state.accelerators.size(); a++) {
    Accelerator &acc = state.accelerators[a];
    if (a % 5 == 0 && !acc.listeners.empty()) acc.listeners[0].from_ports.push_back(8443);
    if (a % 7 == 0) acc.enabled = false;
    if (a % 11 == 0 && !acc.listeners.empty()) acc.listeners[0].endpoint_groups.clear();
    if (a % 13 == 0) acc.listeners.clear();
    if (a % 6 == 1 && !acc.listeners.empty() && !acc.listeners[0].endpoint_groups.empty()) acc.listeners[0].endpoint_groups[0].endpoint_ids = {"arn:stale"};
  }
  for (auto &z : state.zones)
    for (size_t r = 0; r < z.records.size(); r++)
      if (z.records[r].has_alias && r % 9 == 1) z.records[r].alias_dns = "stale.awsglobalaccelerator.com.";
  for (int round = 0; round < 5; round++) {
    Round r = run_round(e, objects, state, cluster);
    report("repair", round, r, state);
    if (r.n_ops == 0) break;
  }
  // phase 3: a third of the objects leave the cache, some lose their annotations -> cleanup
  std::vector<KObject> kept;
  for (size_t i = 0; i < objects.size(); i++) {
    if (i % 3 == 2) continue;
    KObject k = objects[i];
    if (i % 8 == 0) {
      std::vector<std::pair<std::string, std::string>> keep;
      for (auto &a : k.annotations)
        if (a.first != ann + "global-accelerator-managed" && a.first != ann + "route53-hostname") keep.push_back(a);
      k.annotations = keep;
    }
    kept.push_back(k);
  }
  for (int round = 0; round < 5; round++) {
    Round r = run_round(e, kept, state, cluster);
    report("cleanup", round, r, state);
    if (r.n_ops == 0) break;
  }
  gar_engine_destroy(e);
  return 0;
}
