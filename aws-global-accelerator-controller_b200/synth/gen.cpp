// gen.cpp — deterministic synthetic snapshots for tests and bench.py (libgarecon_synth.so).
//
// Produces the SoA tables of include/garecon.h for the configurations of BASELINE.json: a cluster of
// Service/Ingress objects shaped like the reference's own fixtures (local_e2e/pkg/fixtures/service.go:10-51,
// ingress.go:15-58, config/samples/*.yaml) and matching AWS-side lists with a controlled mix of in-sync,
// missing, drifted and orphaned resources (SURVEY.md §8.4 row d).  Every object's fate is a pure function of
// (seed, object index); the AWS-side lists are written in a pseudo-random permutation of object order, as
// ListAccelerators / DescribeLoadBalancers order is unrelated to informer order.
//
// This is workload generation, not product logic: nothing here decides anything about the diff.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "gsyn.h"

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
};
uint64_t mix(uint64_t a, uint64_t b) {
  Rng r(a * 0x9E3779B97F4A7C15ull + b);
  return r.next();
}

struct Slab {
  std::string b;
  std::unordered_map<std::string, gar_str> interned;
  // constant strings (annotation keys, tag keys): one copy in the slab, shared by every row that references it
  gar_str intern(const char *p, size_t n, bool on) {
    if (!on) return put(p, n);
    auto it = interned.find(std::string(p, n));
    if (it != interned.end()) return it->second;
    gar_str r = put(p, n);
    interned.emplace(std::string(p, n), r);
    return r;
  }
  gar_str put(const char *p, size_t n) {
    uint64_t off = b.size();
    b.append(p, n);
    return GAR_STR(off, n);
  }
  gar_str put(const std::string &s) { return put(s.data(), s.size()); }
};

void hex(std::string &out, uint64_t v, int digits) {
  static const char *d = "0123456789abcdef";
  for (int i = digits - 1; i >= 0; i--) out.push_back(d[(v >> (4 * i)) & 15]);
}
std::string hexs(uint64_t v, int digits) {
  std::string s;
  hex(s, v, digits);
  return s;
}
std::string num(uint64_t v, int width = 0) {
  char buf[32];
  snprintf(buf, sizeof(buf), "%0*llu", width, (unsigned long long)v);
  return buf;
}

const char *kRegions[8] = {"us-east-1", "us-west-2", "eu-west-1", "eu-central-1", "ap-northeast-1", "ap-southeast-2", "sa-east-1", "ca-central-1"};
#define ANNP "aws-global-accelerator-controller.h3poteto.dev/"

}  // namespace

extern "C" {

/* struct gsyn_config: see gsyn.h */

// presets follow BASELINE.json "configs"
void gsyn_preset(int cfg, uint32_t n, gsyn_config *c) {
  memset(c, 0, sizeof(*c));
  c->seed = 0x5EEDull + (uint64_t)cfg;
  c->n_objects = n;
  strcpy(c->cluster, "default");
  c->frac_ingress = 0.5f;
  c->n_zones = n / 1000 < 8 ? 8 : (n / 1000 > 1000 ? 1000 : n / 1000);
  c->frac_r53 = 0.6f;
  c->min_hostnames = 1;
  c->max_hostnames = 1;
  c->frac_wildcard = 0.02f;
  c->svc_ports = 2;
  c->hot_pool = 16;
  c->frac_listen_ann = 0.7f;
  c->frac_unmanaged = 0.01f;
  c->frac_ineligible = 0.02f;
  c->p_missing_acc = 0.10f;
  c->p_port_drift = 0.05f;
  c->p_proto_drift = 0.03f;
  c->p_tag_drift = 0.03f;
  c->p_missing_listener = 0.02f;
  c->p_missing_eg = 0.02f;
  c->p_lb_not_active = 0.02f;
  c->p_orphan_acc = 0.03f;
  c->p_rec_missing = 0.15f;
  c->p_alias_drift = 0.05f;
  c->p_orphan_rec = 0.05f;
  c->intern_keys = 1;
  switch (cfg) {
    case 1:  // 100 Services type LoadBalancer, fixture shape (plumbing)
      c->frac_ingress = 0.0f;
      c->frac_r53 = 1.0f;
      c->frac_unmanaged = 0;
      c->frac_ineligible = 0;
      break;
    case 2:  // 10^5 Service+Ingress vs mocked AWS lists
      break;
    case 3:  // 10^6 with multi-hostname route53 annotation (roofline capture config)
      c->frac_r53 = 1.0f;
      c->min_hostnames = 2;
      c->max_hostnames = 4;
      c->n_zones = 1000;
      break;
    case 4:  // 10^7, cfg2 u cfg3
      c->frac_r53 = 0.8f;
      c->min_hostnames = 1;
      c->max_hostnames = 4;
      c->n_zones = 1000;
      break;
    case 5:  // adversarial: colliding hostnames + 64-port listeners
      c->frac_r53 = 1.0f;
      c->min_hostnames = 2;
      c->max_hostnames = 4;
      c->n_zones = 1000;
      c->frac_hot = 0.9f;
      c->svc_ports = 64;
      c->p_port_drift = 0.10f;
      c->p_dup_ports = 0.3f;
      break;
  }
}

}  // extern "C"

namespace {

enum Scenario { SYNC, MISSING_ACC, PORT_DRIFT, PROTO_DRIFT, TAG_DRIFT, MISSING_LISTENER, MISSING_EG, LB_NOT_ACTIVE };

struct ObjSpec {
  uint32_t i;
  bool ingress, eligible, managed, internal_alb, has_listen_ann, udp, has_user_tags, dup_ports;
  std::string ns, name, lb_name, region, hostname, lb_arn, owner, acc_name;
  Scenario sc;
  uint32_t n_host;
  uint64_t r;  // per-object random stream seed
};

struct Gen {
  gsyn_config c;
  explicit Gen(const gsyn_config &cfg) : c(cfg) {}

  ObjSpec spec(uint32_t i) const {
    ObjSpec s;
    const uint32_t gi = c.index_base + i;  // cluster-wide index: names and random streams depend on it, not on the chunk
    s.i = gi;
    Rng r(mix(c.seed, gi));
    s.r = r.next();
    s.ingress = r.uni() < c.frac_ingress;
    s.eligible = r.uni() >= c.frac_ineligible;
    s.managed = r.uni() >= c.frac_unmanaged;
    s.ns = "team-" + num(r.below(50), 2);
    s.name = (s.ingress ? "ing-" : "svc-") + num(gi, 7);
    s.region = kRegions[r.below(8)];
    s.internal_alb = s.ingress && r.uni() < 0.3;
    uint64_t h1 = r.next(), h2 = r.next();
    if (s.ingress) {
      s.lb_name = "k8s-" + s.ns + "-" + s.name + "-" + hexs(h1, 10);
      s.hostname = (s.internal_alb ? "internal-" : "") + s.lb_name + "-" + num(100000000 + h2 % 900000000) + "." + s.region + ".elb.amazonaws.com";
      s.lb_arn = "arn:aws:elasticloadbalancing:" + s.region + ":123456789012:loadbalancer/app/" + s.lb_name + "/" + hexs(h2, 16);
    } else {
      s.lb_name = hexs(h1, 16) + hexs(h2, 16);
      s.hostname = s.lb_name + "-" + hexs(mix(h1, h2), 16) + ".elb." + s.region + ".amazonaws.com";
      s.lb_arn = "arn:aws:elasticloadbalancing:" + s.region + ":123456789012:loadbalancer/net/" + s.lb_name + "/" + hexs(mix(h2, h1), 16);
    }
    s.owner = std::string(s.ingress ? "ingress/" : "service/") + s.ns + "/" + s.name;
    s.acc_name = std::string(s.ingress ? "ingress-" : "service-") + s.ns + "-" + s.name;
    s.has_listen_ann = s.ingress && r.uni() < c.frac_listen_ann;
    s.udp = !s.ingress && r.uni() < 0.1;
    s.has_user_tags = r.uni() < 0.1;
    s.dup_ports = !s.ingress && r.uni() < c.p_dup_ports;
    double u = r.uni();
    double acc = 0;
    s.sc = SYNC;
    const float probs[7] = {c.p_missing_acc, c.p_port_drift, c.p_proto_drift, c.p_tag_drift, c.p_missing_listener, c.p_missing_eg, c.p_lb_not_active};
    for (int k = 0; k < 7; k++) {
      acc += probs[k];
      if (u < acc) {
        s.sc = (Scenario)(k + 1);
        break;
      }
    }
    bool r53 = r.uni() < c.frac_r53;
    s.n_host = r53 ? c.min_hostnames + r.below(c.max_hostnames - c.min_hostnames + 1) : 0;
    return s;
  }
  // desired listener ports of an object
  std::vector<int32_t> ports(const ObjSpec &s) const {
    std::vector<int32_t> p;
    if (s.ingress) {
      if (s.has_listen_ann) return {443};
      return {80};
    }
    if (c.svc_ports <= 2) {
      p = {80, 443};
      p.resize(c.svc_ports);
    } else {
      Rng r(mix(s.r, 77));
      for (uint32_t k = 0; k < c.svc_ports; k++) p.push_back(1000 + (int32_t)k * 7 + (int32_t)r.below(5));
    }
    if (s.dup_ports && p.size() >= 2) p[p.size() - 1] = p[0];
    return p;
  }
  struct Host {
    std::string name;  // as written in the annotation
    uint32_t zone;
    bool hot;
    int fate;  // 0 in sync, 1 missing, 2 alias drift
  };
  // z = zone row of THIS chunk; the name carries the cluster-wide zone row
  std::string zone_name(uint32_t z) const { return "z" + num(c.zone_base + z, 4) + ".example" + num((c.zone_base + z) % 7) + ".com"; }
  Host host(const ObjSpec &s, uint32_t k) const {
    Rng r(mix(s.r, 1000 + k));
    Host h;
    h.hot = r.uni() < c.frac_hot;
    if (h.hot) {
      uint32_t p = r.below(c.hot_pool);
      h.zone = p % c.n_zones;
      h.name = "hot" + num(p, 2) + "." + zone_name(h.zone);
    } else {
      h.zone = r.below(c.n_zones);
      bool wild = r.uni() < c.frac_wildcard;
      h.name = std::string(wild ? "*." : "") + "h" + num(s.i, 7) + "-" + num(k) + (r.uni() < 0.3 ? ".apps." : ".") + zone_name(h.zone);
    }
    double u = r.uni();
    h.fate = u < c.p_rec_missing ? 1 : (u < c.p_rec_missing + c.p_alias_drift ? 2 : 0);
    return h;
  }
  std::string acc_dns(uint32_t i) const { return "a" + hexs(mix(c.seed ^ 0xACCull, i), 16) + ".awsglobalaccelerator.com"; }
  std::string owner_value(const ObjSpec &s) const {
    return "\"heritage=aws-global-accelerator-controller,cluster=" + std::string(c.cluster) + "," + s.owner + "\"";
  }
};

uint32_t gcd(uint32_t a, uint32_t b) { return b ? gcd(b, a % b) : a; }
struct Perm {  // i -> (i*a + b) mod n, a coprime to n
  uint32_t n, a, b;
  Perm(uint32_t n_, uint64_t seed) : n(n_ ? n_ : 1) {
    a = (uint32_t)(mix(seed, 1) % n) | 1u;
    while (gcd(a, n) != 1) a += 2;
    if (n == 1) a = 1;
    b = (uint32_t)(mix(seed, 2) % n);
  }
  uint32_t operator()(uint32_t i) const { return (uint32_t)(((uint64_t)i * a + b) % n); }
};

struct Snapshot {
  gar_objects o{};
  gar_actual a{};
  Slab os, as;
  std::vector<uint8_t> obj_kind, obj_spec, obj_flags, lb_state, acc_enabled, lis_proto, rec_type, rec_has_alias;
  std::vector<gar_str> obj_ns, obj_name, obj_icls, ann_key, ann_val, lbi_host, port_proto;
  std::vector<uint32_t> ann_b, lbi_b, port_b;
  std::vector<int32_t> port_num, pr_from;
  std::vector<gar_str> lb_region, lb_name, lb_dns, lb_arn, acc_arn, acc_name, acc_dns, tag_key, tag_val, lis_arn, eg_arn, ep_id;
  std::vector<uint32_t> tag_b, lis_b, pr_b, eg_b, ep_b;
  std::vector<gar_str> zone_id, zone_name, rec_name, rec_alias, val_value;
  std::vector<uint32_t> rec_b, val_b;
};

void add_accel(Snapshot &S, const Gen &G, const std::string &arn_id, const std::string &name, const std::string &dns, bool enabled, const std::string &owner,
               const std::string &thost, const std::string &cluster, const std::vector<std::pair<std::string, std::string>> &user_tags, int n_lis,
               const std::vector<int32_t> &ports, bool udp, int n_eg, const std::string &endpoint) {
  std::string arn = "arn:aws:globalaccelerator::123456789012:accelerator/" + arn_id;
  S.acc_arn.push_back(0);  // ARNs stay on the host side: the diff never reads them
  S.acc_name.push_back(S.as.put(name));
  S.acc_dns.push_back(S.as.put(dns));
  S.acc_enabled.push_back(enabled ? 1 : 0);
  auto tag = [&](const char *k, const std::string &v) {
    S.tag_key.push_back(S.as.intern(k, strlen(k), G.c.intern_keys != 0));
    S.tag_val.push_back(S.as.put(v));
  };
  tag("aws-global-accelerator-controller-managed", "true");
  tag("aws-global-accelerator-owner", owner);
  tag("aws-global-accelerator-target-hostname", thost);
  tag("aws-global-accelerator-cluster", cluster);
  for (auto &t : user_tags) {
    S.tag_key.push_back(S.as.intern(t.first.data(), t.first.size(), G.c.intern_keys != 0));
    S.tag_val.push_back(S.as.put(t.second));
  }
  S.tag_b.push_back((uint32_t)S.tag_key.size());
  for (int l = 0; l < n_lis; l++) {
    std::string larn = arn + "/listener/" + hexs(mix(S.lis_arn.size(), 5), 8);
    S.lis_arn.push_back(0);
    S.lis_proto.push_back(udp ? GAR_PROTO_UDP : GAR_PROTO_TCP);
    for (int32_t p : ports) S.pr_from.push_back(p);
    S.pr_b.push_back((uint32_t)S.pr_from.size());
    for (int e = 0; e < n_eg; e++) {
      S.eg_arn.push_back(0);
      S.ep_id.push_back(S.as.put(endpoint));
      S.ep_b.push_back((uint32_t)S.ep_id.size());
    }
    S.eg_b.push_back((uint32_t)S.eg_arn.size());
  }
  S.lis_b.push_back((uint32_t)S.lis_arn.size());
}

std::string uuid(uint64_t a) { return hexs(a, 8) + "-" + hexs(a >> 32, 4) + "-" + hexs(mix(a, 1), 4) + "-" + hexs(mix(a, 2), 4) + "-" + hexs(mix(a, 3), 12); }

Snapshot *generate(const gsyn_config &cfg) {
  Gen G(cfg);
  auto *Sp = new Snapshot();
  Snapshot &S = *Sp;
  const uint32_t n = cfg.n_objects;
  S.os.b.reserve((size_t)n * 600 + 1024);
  S.as.b.reserve((size_t)n * 1100 + 1024);
  S.ann_b.push_back(0);
  S.lbi_b.push_back(0);
  S.port_b.push_back(0);
  auto ann = [&](const char *k, const std::string &v) {
    S.ann_key.push_back(S.os.intern(k, strlen(k), cfg.intern_keys != 0));
    S.ann_val.push_back(S.os.put(v));
  };
  gar_str tcp = 0, udp = 0;
  // every object's spec, computed once (in parallel: it is a pure function of (seed, i))
  std::vector<ObjSpec> specs(n);
  {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
      th.emplace_back([&, t] {
        for (uint64_t i = (uint64_t)n * t / nt; i < (uint64_t)n * (t + 1) / nt; i++) specs[i] = G.spec((uint32_t)i);
      });
    for (auto &x : th) x.join();
  }
  // emit_mask (0 = everything): chunks of a sharded cluster only need some table families of a chunk
  const uint32_t mask = cfg.emit_mask ? cfg.emit_mask : 0xFu;
  const uint32_t n_obj = (mask & 1u) ? n : 0, n_lb = (mask & 2u) ? n : 0, n_acc = (mask & 4u) ? n : 0, n_rec = (mask & 8u) ? n : 0;
  // ---------------- objects (informer cache order)
  for (uint32_t i = 0; i < n_obj; i++) {
    const ObjSpec &s = specs[i];
    S.obj_kind.push_back(s.ingress ? GAR_KIND_INGRESS : GAR_KIND_SERVICE);
    S.obj_spec.push_back(s.ingress ? 0 : (s.eligible ? GAR_SVC_LOADBALANCER : GAR_SVC_CLUSTERIP));
    S.obj_flags.push_back(s.ingress ? GAR_OBJ_HAS_INGRESS_CLASS : 0);
    gar_str key = S.os.put(s.ns + "/" + s.name);
    S.obj_ns.push_back(GAR_STR(GAR_STR_OFF(key), s.ns.size()));
    S.obj_name.push_back(GAR_STR(GAR_STR_OFF(key) + s.ns.size() + 1, s.name.size()));
    S.obj_icls.push_back(s.ingress ? S.os.put(s.eligible ? "alb" : "nginx") : 0);
    if (s.managed) ann(ANNP "global-accelerator-managed", "true");
    if (s.n_host) {
      std::string hs;
      for (uint32_t k = 0; k < s.n_host; k++) {
        if (k) hs += ",";
        hs += G.host(s, k).name;
      }
      ann(ANNP "route53-hostname", hs);
    }
    if (s.has_user_tags) ann(ANNP "global-accelerator-tags", "Environment=prod,Team=" + s.ns);
    if (s.ingress) {
      ann("alb.ingress.kubernetes.io/scheme", s.internal_alb ? "internal" : "internet-facing");
      ann("alb.ingress.kubernetes.io/certificate-arn", "arn:aws:acm:" + s.region + ":123456789012:certificate/" + uuid(s.r));
      if (s.has_listen_ann) ann("alb.ingress.kubernetes.io/listen-ports", "[{\"HTTPS\":443}]");
    } else {
      ann("service.beta.kubernetes.io/aws-load-balancer-backend-protocol", "tcp");
      ann("service.beta.kubernetes.io/aws-load-balancer-cross-zone-load-balancing-enabled", "true");
      ann("service.beta.kubernetes.io/aws-load-balancer-type", "nlb");
      ann("service.beta.kubernetes.io/aws-load-balancer-scheme", "internet-facing");
    }
    S.ann_b.push_back((uint32_t)S.ann_key.size());
    S.lbi_host.push_back(S.os.put(s.hostname));
    S.lbi_b.push_back((uint32_t)S.lbi_host.size());
    if (s.ingress) {
      S.port_num.push_back(80);
      S.port_proto.push_back(0);
    } else {
      for (int32_t p : G.ports(s)) {
        S.port_num.push_back(p);
        if (s.udp) {
          if (!udp) udp = S.os.put("UDP");
          S.port_proto.push_back(udp);
        } else {
          if (!tcp) tcp = S.os.put("TCP");
          S.port_proto.push_back(tcp);
        }
      }
    }
    S.port_b.push_back((uint32_t)S.port_num.size());
  }
  // ---------------- load balancers (DescribeLoadBalancers order: permuted)
  {
    Perm pi(n, cfg.seed ^ 0x1B);
    for (uint32_t r = 0; r < n_lb; r++) {
      const ObjSpec &s = specs[pi(r)];
      S.lb_region.push_back(S.as.put(s.region));
      S.lb_name.push_back(S.as.put(s.lb_name));
      S.lb_dns.push_back(S.as.put(s.hostname));
      S.lb_arn.push_back(S.as.put(s.lb_arn));
      S.lb_state.push_back(s.sc == LB_NOT_ACTIVE ? GAR_LB_PROVISIONING : GAR_LB_ACTIVE);
    }
  }
  // ---------------- accelerators (ListAccelerators order: permuted) + orphans interleaved
  S.tag_b.push_back(0);
  S.lis_b.push_back(0);
  S.pr_b.push_back(0);
  S.eg_b.push_back(0);
  S.ep_b.push_back(0);
  {
    Perm pi(n, cfg.seed ^ 0xACC);
    for (uint32_t r = 0; r < n_acc; r++) {
      uint32_t i = pi(r);
      const ObjSpec &s = specs[i];
      Rng rr(mix(s.r, 31));
      bool has = s.eligible && s.sc != MISSING_ACC;
      if (!s.managed) has = rr.uni() < 0.5;  // unmanaged objects sometimes still own one: the cleanup path
      if (!s.eligible) has = false;
      if (has) {
        std::vector<int32_t> ports = G.ports(s);
        bool lis_udp = s.udp;
        std::vector<std::pair<std::string, std::string>> ut;
        if (s.has_user_tags) ut = {{"Environment", "prod"}, {"Team", s.ns}};
        std::string name = s.acc_name;
        int n_lis = 1, n_eg = 1;
        std::string endpoint = s.lb_arn, thost = s.hostname;
        bool enabled = true;
        switch (s.sc) {
          case PORT_DRIFT:
            if (rr.uni() < 0.5 || ports.size() < 2) ports.push_back(8443);
            else ports[rr.below((uint32_t)ports.size())] = 9443;
            break;
          case PROTO_DRIFT: lis_udp = !lis_udp; break;
          case TAG_DRIFT: {
            double v = rr.uni();
            if (v < 0.4) name = "renamed-" + name;
            else if (v < 0.7) thost = "stale." + s.hostname;
            else if (s.has_user_tags) ut.pop_back();
            else enabled = false;
            break;
          }
          case MISSING_LISTENER: n_lis = 0; break;
          case MISSING_EG:
            if (rr.uni() < 0.5) n_eg = 0;
            else endpoint = "arn:aws:elasticloadbalancing:" + s.region + ":123456789012:loadbalancer/net/replaced/" + hexs(s.r, 16);
            break;
          default: break;
        }
        add_accel(S, G, uuid(mix(s.r, 41)), name, G.acc_dns(cfg.index_base + i), enabled, s.owner, thost, cfg.cluster, ut, n_lis, ports, lis_udp, n_eg, endpoint);
      }
      if (rr.uni() < cfg.p_orphan_acc) {  // accelerator whose owner left the cache
        std::string okind = rr.uni() < 0.5 ? "service" : "ingress";
        std::string gone = okind + "/team-" + num(rr.below(50), 2) + "/gone-" + num(cfg.index_base + i, 7);
        add_accel(S, G, uuid(mix(s.r, 43)), okind + "-gone-" + num(cfg.index_base + i, 7), "a" + hexs(mix(s.r, 44), 16) + ".awsglobalaccelerator.com", true, gone,
                  "gone-" + hexs(s.r, 16) + ".elb.us-east-1.amazonaws.com", rr.uni() < 0.9 ? cfg.cluster : "other-cluster", {}, 1, {80, 443}, false, 1,
                  "arn:aws:elasticloadbalancing:us-east-1:123456789012:loadbalancer/net/gone/" + hexs(s.r, 16));
      }
    }
  }
  // ---------------- route53: zones with their record sets
  {
    const uint32_t nz = cfg.n_zones;
    struct Rec {
      std::string name;
      uint8_t type;
      bool alias;
      std::string alias_dns;
      std::vector<std::string> values;
    };
    std::vector<std::vector<Rec>> zr(nz);
    // hot names: one TXT set (values appended per claimant) + one alias A set
    std::vector<int32_t> hot_txt(cfg.hot_pool ? cfg.hot_pool : 1, -1);
    auto esc = [](const std::string &h) {
      std::string o = h;
      size_t p = o.find('*');
      if (p != std::string::npos) o.replace(p, 1, "\\052");
      return o + ".";
    };
    Perm pi(n, cfg.seed ^ 0x53);
    for (uint32_t r = 0; r < n_rec; r++) {
      uint32_t i = pi(r);
      const ObjSpec &s = specs[i];
      if (!s.eligible && !s.ingress) continue;
      std::string ov = G.owner_value(s);
      std::string adns = G.acc_dns(cfg.index_base + i) + ".";
      for (uint32_t k = 0; k < s.n_host; k++) {
        Gen::Host h = G.host(s, k);
        if (h.fate == 1) continue;
        if (h.hot) {
          uint32_t p = (uint32_t)atoi(h.name.c_str() + 3);
          if (hot_txt[p] < 0) {
            hot_txt[p] = (int32_t)zr[h.zone].size();
            zr[h.zone].push_back({esc(h.name), GAR_RR_TXT, false, "", {}});
            zr[h.zone].push_back({esc(h.name), GAR_RR_A, true, adns, {}});
          }
          zr[h.zone][hot_txt[p]].values.push_back(ov);
          continue;
        }
        zr[h.zone].push_back({esc(h.name), GAR_RR_TXT, false, "", {ov}});
        zr[h.zone].push_back({esc(h.name), GAR_RR_A, true, h.fate == 2 ? "stale.awsglobalaccelerator.com." : adns, {}});
      }
      Rng rr(mix(s.r, 57));
      if (rr.uni() < cfg.p_orphan_rec) {  // records of an owner that left the cache
        uint32_t z = rr.below(nz);
        std::string okind = rr.uni() < 0.5 ? "service" : "ingress";
        std::string nm = "gone" + num(cfg.index_base + i, 7) + "." + G.zone_name(z) + ".";
        std::string gv = "\"heritage=aws-global-accelerator-controller,cluster=" + std::string(rr.uni() < 0.9 ? cfg.cluster : "other") + "," + okind + "/team-" +
                         num(rr.below(50), 2) + "/gone-" + num(cfg.index_base + i, 7) + "\"";
        zr[z].push_back({nm, GAR_RR_TXT, false, "", {gv}});
        zr[z].push_back({nm, GAR_RR_A, true, "a" + hexs(mix(s.r, 58), 16) + ".awsglobalaccelerator.com.", {}});
      }
      if (rr.uni() < 0.05) {  // unrelated records
        uint32_t z = rr.below(nz);
        zr[z].push_back({"www" + num(cfg.index_base + i, 7) + "." + G.zone_name(z) + ".", GAR_RR_CNAME, false, "", {"target." + G.zone_name(z) + "."}});
      }
    }
    S.rec_b.push_back(0);
    S.val_b.push_back(0);
    // sharded clusters (zones_total > 0): the zone table is the cluster-wide one on every chunk, record sets only under
    // this chunk's zones [zone_base, zone_base + n_zones)
    const uint32_t zt = cfg.zones_total ? cfg.zones_total : nz, zb = cfg.zones_total ? cfg.zone_base : 0;
    for (uint32_t gz = 0; gz < zt; gz++) {
      const bool mine = gz >= zb && gz < zb + nz;
      const uint32_t z = mine ? gz - zb : 0;
      S.zone_id.push_back(0);
      S.zone_name.push_back(S.as.put("z" + num(gz, 4) + ".example" + num(gz % 7) + ".com."));
      if (!mine) {
        S.rec_b.push_back((uint32_t)S.rec_name.size());
        continue;
      }
      for (auto &rc : zr[z]) {
        S.rec_name.push_back(S.as.put(rc.name));
        S.rec_type.push_back(rc.type);
        S.rec_has_alias.push_back(rc.alias ? 1 : 0);
        S.rec_alias.push_back(rc.alias ? S.as.put(rc.alias_dns) : 0);
        for (auto &v : rc.values) S.val_value.push_back(S.as.put(v));
        S.val_b.push_back((uint32_t)S.val_value.size());
      }
      S.rec_b.push_back((uint32_t)S.rec_name.size());
      std::vector<Rec>().swap(zr[z]);
    }
  }
  // ---------------- optional re-layout: column-major slabs (every column's strings contiguous, in row order)
  if (cfg.layout == 1) {
    auto relayout = [](std::string &slab, std::vector<std::vector<gar_str> *> cols, std::vector<gar_str> *ns, std::vector<gar_str> *nm) {
      std::string out;
      out.reserve(slab.size());
      if (ns) {  // "ns/name" is one string: obj_ns and obj_name stay slices of it
        for (size_t i = 0; i < ns->size(); i++) {
          uint64_t off = GAR_STR_OFF((*ns)[i]), nl = GAR_STR_LEN((*ns)[i]), ml = GAR_STR_LEN((*nm)[i]);
          uint64_t no = out.size();
          out.append(slab, off, nl + 1 + ml);
          (*ns)[i] = GAR_STR(no, nl);
          (*nm)[i] = GAR_STR(no + nl + 1, ml);
        }
      }
      for (auto *col : cols) {
        std::unordered_map<uint64_t, gar_str> seen;  // interned strings (one ref shared by many rows) are copied once
        for (auto &r : *col) {
          if (GAR_STR_LEN(r) == 0) {
            r = 0;
            continue;
          }
          auto it = seen.find(r);
          if (it != seen.end()) {
            r = it->second;
            continue;
          }
          gar_str nr = GAR_STR(out.size(), GAR_STR_LEN(r));
          out.append(slab, GAR_STR_OFF(r), GAR_STR_LEN(r));
          if (seen.size() < 4096) seen.emplace(r, nr);
          r = nr;
        }
      }
      slab.swap(out);
    };
    relayout(S.os.b, {&S.obj_icls, &S.ann_key, &S.ann_val, &S.lbi_host, &S.port_proto}, &S.obj_ns, &S.obj_name);
    relayout(S.as.b, {&S.lb_region, &S.lb_name, &S.lb_dns, &S.lb_arn, &S.acc_name, &S.acc_dns, &S.tag_key, &S.tag_val, &S.ep_id, &S.zone_name, &S.rec_name,
                      &S.rec_alias, &S.val_value},
             nullptr, nullptr);
  }
  // ---------------- publish
  S.o.slab_len = S.os.b.size();
  S.a.slab_len = S.as.b.size();
  S.os.b.append(64, '\0');
  S.as.b.append(64, '\0');
  gar_objects &o = S.o;
  o.n_objects = n_obj;
  o.obj_kind = S.obj_kind.data();
  o.obj_spec_type = S.obj_spec.data();
  o.obj_flags = S.obj_flags.data();
  o.obj_ns = S.obj_ns.data();
  o.obj_name = S.obj_name.data();
  o.obj_ingress_class = S.obj_icls.data();
  o.obj_ann_begin = S.ann_b.data();
  o.obj_lbi_begin = S.lbi_b.data();
  o.obj_port_begin = S.port_b.data();
  o.n_ann = (uint32_t)S.ann_key.size();
  o.ann_key = S.ann_key.data();
  o.ann_val = S.ann_val.data();
  o.n_lbi = (uint32_t)S.lbi_host.size();
  o.lbi_hostname = S.lbi_host.data();
  o.n_ports = (uint32_t)S.port_num.size();
  o.port_number = S.port_num.data();
  o.port_proto = S.port_proto.data();
  o.slab = (const uint8_t *)S.os.b.data();
  gar_actual &a = S.a;
  a.n_lbs = (uint32_t)S.lb_name.size();
  a.lb_region = S.lb_region.data();
  a.lb_name = S.lb_name.data();
  a.lb_dns = S.lb_dns.data();
  a.lb_arn = S.lb_arn.data();
  a.lb_state = S.lb_state.data();
  a.n_accels = (uint32_t)S.acc_arn.size();
  a.acc_name = S.acc_name.data();
  a.acc_dns = S.acc_dns.data();
  a.acc_enabled = S.acc_enabled.data();
  a.acc_tag_begin = S.tag_b.data();
  a.acc_lis_begin = S.lis_b.data();
  a.n_tags = (uint32_t)S.tag_key.size();
  a.tag_key = S.tag_key.data();
  a.tag_val = S.tag_val.data();
  a.n_listeners = (uint32_t)S.lis_arn.size();
  a.lis_proto = S.lis_proto.data();
  a.lis_pr_begin = S.pr_b.data();
  a.lis_eg_begin = S.eg_b.data();
  a.n_port_ranges = (uint32_t)S.pr_from.size();
  a.pr_from = S.pr_from.data();
  a.n_egs = (uint32_t)S.eg_arn.size();
  a.eg_ep_begin = S.ep_b.data();
  a.n_endpoints = (uint32_t)S.ep_id.size();
  a.ep_id = S.ep_id.data();
  a.n_zones = (uint32_t)S.zone_name.size();
  a.zone_name = S.zone_name.data();
  a.zone_rec_begin = S.rec_b.data();
  a.n_records = (uint32_t)S.rec_name.size();
  a.rec_name = S.rec_name.data();
  a.rec_type = S.rec_type.data();
  a.rec_has_alias = S.rec_has_alias.data();
  a.rec_alias_dns = S.rec_alias.data();
  a.rec_val_begin = S.val_b.data();
  a.n_values = (uint32_t)S.val_value.size();
  a.val_value = S.val_value.data();
  a.slab = (const uint8_t *)S.as.b.data();
  return Sp;
}

}  // namespace

extern "C" {

// Generate a snapshot; the returned handle owns every buffer the two table structs point to.
gsyn_snapshot *gsyn_generate(const gsyn_config *cfg) { return (gsyn_snapshot *)generate(*cfg); }
const gar_objects *gsyn_objects(const gsyn_snapshot *s) { return &((const Snapshot *)s)->o; }
const gar_actual *gsyn_actual(const gsyn_snapshot *s) { return &((const Snapshot *)s)->a; }
void gsyn_free(gsyn_snapshot *s) { delete (Snapshot *)s; }

}  // extern "C"
