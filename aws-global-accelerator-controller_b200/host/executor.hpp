// executor.hpp — host side of the C ABI, part 2: change set -> the reference's side effects.
//
// Every op stands for mutations the reference performs in-line (include/garecon.h op table).  The executor walks the ops in
// order and calls a `Cloud` whose methods are the reference's SDK wrappers (pkg/cloudprovider/aws/global_accelerator.go:654-1013,
// route53.go:183-315); the arguments are derived from the object exactly as the reference derives them
// (acceleratorName :53-60, acceleratorTags :35-51, listenerForService/Ingress :503-557 via the engine's derived bits,
// Route53OwnerValue route53.go:18-20).  `MockCloud` applies the calls to an in-memory CloudState — the stand-in for the mock
// cloudprovider BASELINE config 1 names and the reference does not have — so tests can run diff -> execute -> diff to a fixed
// point.  C++ because this image has no Go toolchain; in Go the Cloud is the reference's *cloudaws.AWS as is.
#pragma once

#include <set>
#include <string>
#include <vector>

#include "packer.hpp"

namespace garecon {

constexpr const char *kAnnPrefix = "aws-global-accelerator-controller.h3poteto.dev/";
constexpr const char *kTagManaged = "aws-global-accelerator-controller-managed";
constexpr const char *kTagOwner = "aws-global-accelerator-owner";
constexpr const char *kTagTargetHostname = "aws-global-accelerator-target-hostname";
constexpr const char *kTagCluster = "aws-global-accelerator-cluster";

// the SDK wrapper surface the ops need (names as in the reference)
struct Cloud {
  virtual ~Cloud() = default;
  // createAccelerator + createListener + createEndpointGroup (global_accelerator.go:213-252, 654-701, 815-837, 971-990)
  virtual void CreateAcceleratorChain(const std::string &name, bool ipv4, const std::vector<std::pair<std::string, std::string>> &tags,
                                      const std::vector<int32_t> &ports, uint8_t proto, bool ip_preserve, const std::string &lb_arn) = 0;
  virtual void UpdateAccelerator(size_t accel, const std::string &name, const std::vector<std::pair<std::string, std::string>> &tags) = 0;  // :703-741
  virtual size_t CreateListener(size_t accel, const std::vector<int32_t> &ports, uint8_t proto) = 0;                                      // :815-837
  virtual void UpdateListener(size_t accel, size_t listener, const std::vector<int32_t> &ports, uint8_t proto) = 0;                        // :839-861
  virtual void CreateEndpointGroup(size_t accel, size_t listener, const std::string &lb_arn, bool ip_preserve) = 0;                        // :971-990
  virtual void UpdateEndpointGroup(size_t accel, size_t listener, size_t eg, const std::string &lb_arn, bool ip_preserve) = 0;             // :992-1010
  virtual void DeleteAcceleratorChain(size_t accel) = 0;                                                                                   // :254-288
  virtual void CreateMetadataAndAliasRecords(size_t zone, const std::string &hostname, const std::string &owner_value, const std::string &accel_dns) = 0;  // route53.go:240-289
  virtual void UpsertAliasRecord(size_t zone, size_t record, const std::string &accel_dns) = 0;                                            // :291-315
  virtual void DeleteRecord(size_t zone, size_t record) = 0;                                                                                // :183-197
};

// row numbering of a packed CloudState (Packer::AddCloud order): flat row -> position in the model
struct RowMap {
  std::vector<std::pair<size_t, size_t>> listener;               // listener row -> (accelerator, listener index)
  std::vector<std::pair<size_t, std::pair<size_t, size_t>>> eg;  // endpoint-group row -> (accelerator, (listener index, eg index))
  std::vector<std::pair<size_t, size_t>> record;                 // record row -> (zone, record index)
  explicit RowMap(const CloudState &c) {
    for (size_t a = 0; a < c.accelerators.size(); a++)
      for (size_t l = 0; l < c.accelerators[a].listeners.size(); l++) {
        listener.push_back({a, l});
        for (size_t g = 0; g < c.accelerators[a].listeners[l].endpoint_groups.size(); g++) eg.push_back({a, {l, g}});
      }
    for (size_t z = 0; z < c.zones.size(); z++)
      for (size_t r = 0; r < c.zones[z].records.size(); r++) record.push_back({z, r});
  }
};

inline const std::string *Annotation(const KObject &k, const std::string &key) {
  for (auto &a : k.annotations)
    if (a.first == key) return &a.second;
  return nullptr;
}
inline std::string ResourceOf(const KObject &k) { return k.kind == GAR_KIND_SERVICE ? "service" : "ingress"; }
// acceleratorName (global_accelerator.go:53-60)
inline std::string AcceleratorName(const KObject &k) {
  const std::string *v = Annotation(k, std::string(kAnnPrefix) + "global-accelerator-name");
  return v && !v->empty() ? *v : ResourceOf(k) + "-" + k.ns + "-" + k.name;
}
// acceleratorTags (:35-51): split on ',' then '='; only pieces with exactly two parts
inline std::vector<std::pair<std::string, std::string>> AcceleratorUserTags(const KObject &k) {
  std::vector<std::pair<std::string, std::string>> out;
  const std::string *v = Annotation(k, std::string(kAnnPrefix) + "global-accelerator-tags");
  if (!v) return out;
  size_t p = 0;
  for (;;) {
    size_t q = v->find(',', p);
    std::string piece = v->substr(p, q == std::string::npos ? std::string::npos : q - p);
    size_t e1 = piece.find('=');
    if (e1 != std::string::npos && piece.find('=', e1 + 1) == std::string::npos) out.push_back({piece.substr(0, e1), piece.substr(e1 + 1)});
    if (q == std::string::npos) break;
    p = q + 1;
  }
  return out;
}
// Route53OwnerValue (route53.go:18-20), quotes included
inline std::string Route53OwnerValue(const std::string &cluster, const KObject &k) {
  return "\"heritage=aws-global-accelerator-controller,cluster=" + cluster + "," + ResourceOf(k) + "/" + k.ns + "/" + k.name + "\"";
}

// Executes one change set.  `objects` / `cloud_state` are what was packed (same order); `cloud` receives the calls.
inline size_t ExecuteChangeSet(const gar_changeset &cs, const std::vector<KObject> &objects, const CloudState &state, const std::string &cluster, Cloud &cloud) {
  RowMap rows(state);
  size_t executed = 0;
  std::vector<std::pair<size_t, size_t>> created_listener;  // (accelerator, listener index) made by GA_CREATE_LISTENER, for the EG op that follows
  for (uint64_t k = 0; k < cs.n_ops; k++) {
    const gar_op &op = cs.ops[k];
    const uint32_t code = op.head & 0xFFu;
    const KObject *ob = op.obj != GAR_NONE ? &objects[op.obj] : nullptr;
    auto desired_ports = [&]() {
      std::vector<int32_t> p;
      if (cs.derived[op.obj] & GAR_DV_PORTS_FROM_ANN)
        for (uint32_t x = cs.dport_begin[op.obj]; x < cs.dport_begin[op.obj + 1]; x++) p.push_back(cs.dports[x]);
      else
        for (auto &pp : ob->ports) p.push_back(pp.first);
      return p;
    };
    auto proto = [&]() -> uint8_t { return (cs.derived[op.obj] & GAR_DV_PROTO_UDP) ? GAR_PROTO_UDP : GAR_PROTO_TCP; };
    auto ip_preserve = [&]() { return (cs.derived[op.obj] & GAR_DV_IP_PRESERVE) != 0; };
    auto system_tags = [&](const LoadBalancer &lb, bool with_cluster) {
      std::vector<std::pair<std::string, std::string>> t = {{kTagManaged, "true"}, {kTagOwner, ResourceOf(*ob) + "/" + ob->ns + "/" + ob->name}, {kTagTargetHostname, lb.dns}};
      if (with_cluster) t.push_back({kTagCluster, cluster});
      for (auto &u : AcceleratorUserTags(*ob)) t.push_back(u);
      return t;
    };
    switch (code) {
      case GAR_OP_GA_CREATE_CHAIN: {
        const LoadBalancer &lb = state.lbs[op.a0];
        cloud.CreateAcceleratorChain(AcceleratorName(*ob), (cs.derived[op.obj] & GAR_DV_IPV4) != 0, system_tags(lb, true), desired_ports(), proto(), ip_preserve(), lb.arn);
        break;
      }
      case GAR_OP_GA_UPDATE_ACCEL: cloud.UpdateAccelerator(op.a0, AcceleratorName(*ob), system_tags(state.lbs[op.a1], false)); break;
      case GAR_OP_GA_CREATE_LISTENER: created_listener.push_back({op.a0, cloud.CreateListener(op.a0, desired_ports(), proto())}); break;
      case GAR_OP_GA_UPDATE_LISTENER: cloud.UpdateListener(op.a0, rows.listener[op.a1].second, desired_ports(), proto()); break;
      case GAR_OP_GA_CREATE_EG: {
        size_t li = op.a1 == GAR_NONE ? created_listener.back().second : rows.listener[op.a1].second;
        cloud.CreateEndpointGroup(op.a0, li, state.lbs[op.a2].arn, ip_preserve());
        break;
      }
      case GAR_OP_GA_UPDATE_EG: cloud.UpdateEndpointGroup(op.a0, rows.eg[op.a1].second.first, rows.eg[op.a1].second.second, state.lbs[op.a2].arn, ip_preserve()); break;
      case GAR_OP_GA_DELETE_CHAIN: cloud.DeleteAcceleratorChain(op.a0); break;
      case GAR_OP_R53_CREATE: {
        const std::string *ann = Annotation(*ob, std::string(kAnnPrefix) + "route53-hostname");
        uint32_t want = op.sub & 0xFFFFFu, idx = 0;  // k-th piece of strings.Split(annotation, ",") (route53/service.go:71)
        size_t p = 0;
        std::string host;
        for (;;) {
          size_t q = ann->find(',', p);
          if (idx == want) {
            host = ann->substr(p, q == std::string::npos ? std::string::npos : q - p);
            break;
          }
          if (q == std::string::npos) break;
          p = q + 1;
          idx++;
        }
        cloud.CreateMetadataAndAliasRecords(op.a0, host, Route53OwnerValue(cluster, *ob), state.accelerators[op.a1].dns);
        break;
      }
      case GAR_OP_R53_UPSERT_A: cloud.UpsertAliasRecord(op.a0, rows.record[op.a2].second, state.accelerators[op.a1].dns); break;
      case GAR_OP_R53_DELETE_RECORD: cloud.DeleteRecord(op.a0, rows.record[op.a1].second); break;
      default: continue;
    }
    executed++;
  }
  return executed;
}

// In-memory AWS.  Calls address existing resources by their position in `before` (the state that was packed); deletions are
// deferred to Commit() so positions stay valid while a change set is executed.
class MockCloud : public Cloud {
 public:
  explicit MockCloud(CloudState *s) : s_(s) {}
  void CreateAcceleratorChain(const std::string &name, bool, const std::vector<std::pair<std::string, std::string>> &tags, const std::vector<int32_t> &ports,
                              uint8_t proto, bool, const std::string &lb_arn) override {
    Accelerator a;
    size_t id = ++serial_;
    a.arn = "arn:aws:globalaccelerator::1:accelerator/mock-" + std::to_string(id);
    a.name = name;
    a.dns = "mock" + std::to_string(id) + ".awsglobalaccelerator.com";
    a.tags = tags;
    Listener l;
    l.arn = a.arn + "/listener/1";
    l.proto = proto;
    l.from_ports = ports;
    l.endpoint_groups.push_back(EndpointGroup{l.arn + "/endpoint-group/1", {lb_arn}});
    a.listeners.push_back(l);
    s_->accelerators.push_back(a);
  }
  void UpdateAccelerator(size_t accel, const std::string &name, const std::vector<std::pair<std::string, std::string>> &tags) override {
    Accelerator &a = s_->accelerators[accel];
    a.enabled = true;
    a.name = name;
    // TagResource: replaces the value of existing keys, appends new ones (later duplicates of `tags` win)
    std::vector<std::pair<std::string, std::string>> fin;
    for (auto &t : tags) {
      bool seen = false;
      for (auto &f : fin)
        if (f.first == t.first) {
          f.second = t.second;
          seen = true;
        }
      if (!seen) fin.push_back(t);
    }
    std::vector<std::pair<std::string, std::string>> out;
    for (auto &t : a.tags) {
      bool replaced = false;
      for (auto &f : fin) replaced = replaced || f.first == t.first;
      if (!replaced) out.push_back(t);
    }
    for (auto &f : fin) out.push_back(f);
    a.tags = out;
  }
  size_t CreateListener(size_t accel, const std::vector<int32_t> &ports, uint8_t proto) override {
    Accelerator &a = s_->accelerators[accel];
    Listener l;
    l.arn = a.arn + "/listener/n" + std::to_string(a.listeners.size());
    l.proto = proto;
    l.from_ports = ports;
    a.listeners.push_back(l);
    return a.listeners.size() - 1;
  }
  void UpdateListener(size_t accel, size_t listener, const std::vector<int32_t> &ports, uint8_t proto) override {
    Listener &l = s_->accelerators[accel].listeners[listener];
    l.from_ports = ports;
    l.proto = proto;
  }
  void CreateEndpointGroup(size_t accel, size_t listener, const std::string &lb_arn, bool) override {
    Listener &l = s_->accelerators[accel].listeners[listener];
    l.endpoint_groups.push_back(EndpointGroup{l.arn + "/endpoint-group/n" + std::to_string(l.endpoint_groups.size()), {lb_arn}});
  }
  void UpdateEndpointGroup(size_t accel, size_t listener, size_t eg, const std::string &lb_arn, bool) override {
    s_->accelerators[accel].listeners[listener].endpoint_groups[eg].endpoint_ids = {lb_arn};
  }
  void DeleteAcceleratorChain(size_t accel) override { dead_acc_.insert(accel); }
  void CreateMetadataAndAliasRecords(size_t zone, const std::string &hostname, const std::string &owner_value, const std::string &accel_dns) override {
    std::string name = hostname;
    size_t star = name.find('*');
    if (star != std::string::npos) name.replace(star, 1, "\\052");
    name += ".";
    RecordSet txt;
    txt.name = name;
    txt.type = GAR_RR_TXT;
    txt.values = {owner_value};
    RecordSet a;
    a.name = name;
    a.type = GAR_RR_A;
    a.has_alias = true;
    a.alias_dns = accel_dns + ".";
    s_->zones[zone].records.push_back(txt);
    s_->zones[zone].records.push_back(a);
  }
  void UpsertAliasRecord(size_t zone, size_t record, const std::string &accel_dns) override {
    RecordSet &r = s_->zones[zone].records[record];
    r.type = GAR_RR_A;
    r.has_alias = true;
    r.alias_dns = accel_dns + ".";
  }
  void DeleteRecord(size_t zone, size_t record) override { dead_rec_.insert({zone, record}); }
  // apply the deferred deletions
  void Commit() {
    for (auto it = dead_acc_.rbegin(); it != dead_acc_.rend(); ++it) s_->accelerators.erase(s_->accelerators.begin() + (long)*it);
    for (auto it = dead_rec_.rbegin(); it != dead_rec_.rend(); ++it) s_->zones[it->first].records.erase(s_->zones[it->first].records.begin() + (long)it->second);
    dead_acc_.clear();
    dead_rec_.clear();
  }

 private:
  CloudState *s_;
  size_t serial_ = 0;
  std::set<size_t> dead_acc_;
  std::set<std::pair<size_t, size_t>> dead_rec_;
};

}  // namespace garecon
