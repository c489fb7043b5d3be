// executor.hpp — host side of the C ABI, part 2: change set -> the reference's side effects.
//
// Every op stands for mutations the reference performs in-line (include/garecon.h op table).  The executor walks the ops in
// order and calls a `Cloud` whose methods are the reference's SDK wrappers, one AWS call each
// (pkg/cloudprovider/aws/global_accelerator.go:654-1013, route53.go:183-315); the arguments are derived from the object exactly
// as the reference derives them (acceleratorName :53-60, acceleratorTags :35-51, listenerForService/Ingress :503-557 via the
// engine's derived bits, Route53OwnerValue route53.go:18-20).  Every call can FAIL, and a failure has the reference's
// consequences:
//   * the first error ends the object's op stream — the remaining ops of that (controller, object) are skipped, the key is
//     requeued rate-limited (globalaccelerator/service.go:107-110, reconcile.go:75-77) — other objects are unaffected;
//   * a create chain that fails after the accelerator exists is rolled back (CleanupGlobalAccelerator on what was created,
//     result ignored: global_accelerator.go:139-147) — EXCEPT the Ingress path, which swallows a listener-create error and
//     reports success with a listener-less accelerator (createGlobalAcceleratorForIngress returns `arn, nil`, :241-244);
//   * Events as the reference records them: GlobalAcceleratorCreated per created chain (service.go:116-118),
//     GlobalAcceleratorDeleted / Route53RecordDeleted after a complete cleanup of an un-annotated object (service.go:82,
//     route53/service.go:67), "Route53RecourdCreated" [sic] per lbIngress whose ensure created a record and returned no error
//     (route53/service.go:101-103; ensureRoute53 drops `created` when a later hostname fails, route53.go:84-123).
// `MockCloud` applies the calls to an in-memory CloudState — the stand-in for the mock cloudprovider BASELINE config 1 names and
// the reference does not have — with fault injection, so tests can run diff -> execute -> diff to a fixed point through
// failures.  C++ because this image has no Go toolchain; in Go the Cloud is the reference's *cloudaws.AWS as is.
#pragma once

#include <functional>
#include <map>
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

struct Status {
  bool ok = true;
  std::string msg;
  static Status OK() { return {}; }
  static Status Err(std::string m) { return {false, std::move(m)}; }
  explicit operator bool() const { return ok; }
};
using Tags = std::vector<std::pair<std::string, std::string>>;

// The SDK wrapper surface the ops need, one AWS call per method (names as in the reference).  Resources are addressed by
// position: existing ones by their position in the state that was packed, created ones by the handle the create call returns.
struct Cloud {
  virtual ~Cloud() = default;
  virtual Status CreateAccelerator(const std::string &name, bool ipv4, const Tags &tags, size_t *accel) = 0;                    // :654-701
  virtual Status UpdateAccelerator(size_t accel, const std::string &name, const Tags &tags) = 0;                                // :703-741 (UpdateAccelerator + TagResource)
  virtual Status CreateListener(size_t accel, const std::vector<int32_t> &ports, uint8_t proto, size_t *listener) = 0;          // :815-837
  virtual Status UpdateListener(size_t accel, size_t listener, const std::vector<int32_t> &ports, uint8_t proto) = 0;           // :839-861
  virtual Status CreateEndpointGroup(size_t accel, size_t listener, const std::string &lb_arn, bool ip_preserve, size_t *eg) = 0;  // :971-990
  virtual Status UpdateEndpointGroup(size_t accel, size_t listener, size_t eg, const std::string &lb_arn, bool ip_preserve) = 0;   // :992-1010
  virtual Status DeleteEndpointGroup(size_t accel, size_t listener, size_t eg) = 0;                                             // :1012-1024
  virtual Status DeleteListener(size_t accel, size_t listener) = 0;                                                             // :863-875
  virtual Status DeleteAccelerator(size_t accel) = 0;                                                                           // :743-787 (disable, wait, delete)
  virtual Status CreateMetadataRecordSet(size_t zone, const std::string &hostname, const std::string &owner_value) = 0;         // route53.go:240-262
  virtual Status CreateRecordSet(size_t zone, const std::string &hostname, const std::string &accel_dns, size_t *record) = 0;   // :264-289
  virtual Status UpdateRecordSet(size_t zone, size_t record, const std::string &accel_dns) = 0;                                 // :291-315
  virtual Status DeleteRecord(size_t zone, size_t record) = 0;                                                                  // :183-197
  virtual std::string AcceleratorArn(size_t accel) const = 0;                                                                   // for the Created event's message
};

// row numbering of a packed CloudState (Packer::AddCloud order): flat row -> position in the model
struct RowMap {
  std::vector<std::pair<size_t, size_t>> listener;               // listener row -> (accelerator, listener index)
  std::vector<std::pair<size_t, std::pair<size_t, size_t>>> eg;  // endpoint-group row -> (accelerator, (listener index, eg index))
  std::vector<std::pair<size_t, size_t>> record;                 // record row -> (zone, record index)
  std::vector<std::pair<size_t, size_t>> value;                  // value row -> (record row, value index)
  explicit RowMap(const CloudState &c) {
    for (size_t a = 0; a < c.accelerators.size(); a++)
      for (size_t l = 0; l < c.accelerators[a].listeners.size(); l++) {
        listener.push_back({a, l});
        for (size_t g = 0; g < c.accelerators[a].listeners[l].endpoint_groups.size(); g++) eg.push_back({a, {l, g}});
      }
    for (size_t z = 0; z < c.zones.size(); z++)
      for (size_t r = 0; r < c.zones[z].records.size(); r++) {
        for (size_t v = 0; v < c.zones[z].records[r].values.size(); v++) value.push_back({record.size(), v});
        record.push_back({z, r});
      }
  }
};

inline const std::string *Annotation(const KObject &k, const std::string &key) {
  for (auto &a : k.annotations)
    if (a.first == key) return &a.second;
  return nullptr;
}
inline std::string ResourceOf(uint32_t kind) { return kind == GAR_KIND_SERVICE ? "service" : "ingress"; }
inline std::string ResourceOf(const KObject &k) { return ResourceOf(k.kind); }
// acceleratorName (global_accelerator.go:53-60)
inline std::string AcceleratorName(const KObject &k) {
  const std::string *v = Annotation(k, std::string(kAnnPrefix) + "global-accelerator-name");
  return v && !v->empty() ? *v : ResourceOf(k) + "-" + k.ns + "-" + k.name;
}
// acceleratorTags (:35-51): split on ',' then '='; only pieces with exactly two parts
inline Tags AcceleratorUserTags(const KObject &k) {
  Tags out;
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
// k-th piece of strings.Split(annotation, ",") (route53/service.go:71)
inline std::string Route53Hostname(const KObject &k, uint32_t want) {
  const std::string *ann = Annotation(k, std::string(kAnnPrefix) + "route53-hostname");
  if (!ann) return "";
  size_t p = 0;
  for (uint32_t idx = 0;; idx++) {
    size_t q = ann->find(',', p);
    if (idx == want) return ann->substr(p, q == std::string::npos ? std::string::npos : q - p);
    if (q == std::string::npos) return "";
    p = q + 1;
  }
}

// c.recorder.Event(obj, corev1.EventTypeNormal, reason, message)
struct Event {
  uint32_t obj;
  uint8_t ctrl;
  std::string type, reason, message;
};
struct ExecReport {
  size_t executed = 0, skipped = 0, failed_calls = 0, rollbacks = 0;
  std::map<std::pair<uint8_t, uint32_t>, std::string> failed;  // (controller, object row) -> first error: the key is requeued rate-limited
  std::map<std::string, std::string> failed_owners;            // "service/ns/name" of a deleted key whose cleanup failed -> error
  std::vector<Event> events;
  bool Failed(uint8_t ctrl, uint32_t obj) const { return failed.count({ctrl, obj}) != 0; }
};
enum ExecFilter : uint32_t { kExecGA = 1u << GAR_CTRL_GA, kExecR53 = 1u << GAR_CTRL_R53, kExecAll = kExecGA | kExecR53 };

// Executes one change set.  `objects` / `state` are what was packed (same order); `cloud` receives the calls.  `filter` selects
// the controller whose ops are executed: the two controllers run separate workers (globalaccelerator/controller.go:208-213,
// route53/controller.go:199-204), each executes its own half of a change set.
inline ExecReport ExecuteChangeSet(const gar_changeset &cs, const std::vector<KObject> &objects, const CloudState &state, const std::string &cluster, Cloud &cloud,
                                   uint32_t filter = kExecAll) {
  RowMap rows(state);
  ExecReport rep;
  // what an object's own earlier ops created (GAR_PENDING arguments, include/garecon.h "Self-observation")
  struct Made {
    size_t accel = SIZE_MAX, listener = SIZE_MAX, eg = SIZE_MAX;
  };
  std::map<uint32_t, Made> made_chain;                                     // object -> its GA_CREATE_CHAIN
  std::map<std::pair<uint32_t, uint32_t>, std::pair<size_t, size_t>> made_eg;  // (object, accelerator row) -> (listener index, eg index) it created
  std::map<std::pair<uint32_t, uint32_t>, size_t> made_listener;           // (object, accelerator row) -> listener index it created
  std::map<std::pair<uint32_t, std::string>, size_t> made_record;          // (object, hostname) -> A record it created
  std::map<std::pair<uint32_t, uint32_t>, bool> r53_created;               // (object, j) -> ensure created a record
  std::set<std::pair<uint32_t, uint32_t>> r53_j_failed;
  auto owner_of_accel = [&](uint32_t acc) {
    std::string v;
    for (auto &t : state.accelerators[acc].tags)
      if (t.first == kTagOwner) v = t.second;
    return v;
  };
  auto owner_of_value = [&](uint32_t vrow) -> std::string {  // "...,cluster=<c>,<resource>/<ns>/<name>" inside the quotes
    auto &rv = rows.value[vrow];
    auto &zr = rows.record[rv.first];
    const std::string &val = state.zones[zr.first].records[zr.second].values[rv.second];
    std::string pre = "\"heritage=aws-global-accelerator-controller,cluster=" + cluster + ",";
    return val.size() > pre.size() + 1 && val.compare(0, pre.size(), pre) == 0 ? val.substr(pre.size(), val.size() - pre.size() - 1) : val;
  };
  for (uint64_t k = 0; k < cs.n_ops; k++) {
    const gar_op &op = cs.ops[k];
    const uint32_t code = op.head & 0xFFu;
    const uint8_t ctrl = (uint8_t)((op.head >> 8) & 0xFFu);
    if (!(filter & (1u << ctrl))) continue;
    const bool orphan = op.obj == GAR_NONE;
    const KObject *ob = orphan ? nullptr : &objects[op.obj];
    std::string owner;
    if (orphan) {
      owner = code == GAR_OP_GA_DELETE_CHAIN ? owner_of_accel(op.a0) : owner_of_value(op.a2);
      if (rep.failed_owners.count(owner)) {  // processDelete of that key already returned its error
        rep.skipped++;
        continue;
      }
    } else if (rep.Failed(ctrl, op.obj)) {  // the object's op stream ended at its first error
      rep.skipped++;
      continue;
    }
    auto fail = [&](const Status &st) {
      rep.failed_calls++;
      if (orphan) rep.failed_owners.emplace(owner, st.msg);
      else rep.failed.emplace(std::make_pair(ctrl, op.obj), st.msg);
      if (ctrl == GAR_CTRL_R53 && !orphan) r53_j_failed.insert({op.obj, op.sub >> 20});
    };
    auto desired_ports = [&]() {
      std::vector<int32_t> p;
      if (cs.derived[op.obj] & GAR_DV_PORTS_FROM_ANN)
        for (uint32_t x = cs.dport_begin[op.obj]; x < cs.dport_begin[op.obj + 1]; x++) p.push_back(cs.dports[x]);
      else
        for (auto &pp : ob->ports) p.push_back(pp.first);
      return p;
    };
    // listenerForService: the derived protocol bit; listenerForIngress: always TCP (global_accelerator.go:522-557)
    auto proto = [&]() -> uint8_t { return (ob->kind == GAR_KIND_SERVICE && (cs.derived[op.obj] & GAR_DV_PROTO_UDP)) ? GAR_PROTO_UDP : GAR_PROTO_TCP; };
    auto ip_preserve = [&]() { return (cs.derived[op.obj] & GAR_DV_IP_PRESERVE) != 0; };
    auto system_tags = [&](const LoadBalancer &lb, bool with_cluster) {
      Tags t = {{kTagManaged, "true"}, {kTagOwner, ResourceOf(*ob) + "/" + ob->ns + "/" + ob->name}, {kTagTargetHostname, lb.dns}};
      if (with_cluster) t.push_back({kTagCluster, cluster});
      for (auto &u : AcceleratorUserTags(*ob)) t.push_back(u);
      return t;
    };
    auto accel_of = [&](uint32_t a) { return a == GAR_PENDING ? made_chain[op.obj].accel : (size_t)a; };
    Status st = Status::OK();
    switch (code) {
      case GAR_OP_GA_CREATE_CHAIN: {  // createGlobalAcceleratorFor{Service,Ingress} (:213-252) + the caller's rollback (:139-147)
        const LoadBalancer &lb = state.lbs[op.a0];
        Made m;
        st = cloud.CreateAccelerator(AcceleratorName(*ob), (cs.derived[op.obj] & GAR_DV_IPV4) != 0, system_tags(lb, true), &m.accel);
        if (!st) break;  // nothing exists yet: nothing to clean up
        st = cloud.CreateListener(m.accel, desired_ports(), proto(), &m.listener);
        if (!st && ob->kind == GAR_KIND_INGRESS) {
          st = Status::OK();  // createGlobalAcceleratorForIngress returns (arn, nil) here: the error is swallowed (:241-244)
          m.listener = SIZE_MAX;
        } else if (st) {
          st = cloud.CreateEndpointGroup(m.accel, m.listener, lb.arn, ip_preserve(), &m.eg);
        }
        if (!st) {  // "some resources are created, so cleanup": CleanupGlobalAccelerator on the new accelerator, result ignored
          rep.rollbacks++;
          if (m.listener != SIZE_MAX) (void)cloud.DeleteListener(m.accel, m.listener);
          (void)cloud.DeleteAccelerator(m.accel);
          break;
        }
        made_chain[op.obj] = m;
        rep.events.push_back({op.obj, ctrl, "Normal", "GlobalAcceleratorCreated", "Global Acclerator is created: " + cloud.AcceleratorArn(m.accel)});
        break;
      }
      case GAR_OP_GA_UPDATE_ACCEL: st = cloud.UpdateAccelerator(accel_of(op.a0), AcceleratorName(*ob), system_tags(state.lbs[op.a1], false)); break;
      case GAR_OP_GA_CREATE_LISTENER: {
        size_t li = SIZE_MAX;
        st = cloud.CreateListener(op.a0, desired_ports(), proto(), &li);
        if (st) made_listener[{op.obj, op.a0}] = li;
        break;
      }
      case GAR_OP_GA_UPDATE_LISTENER: st = cloud.UpdateListener(op.a0, rows.listener[op.a1].second, desired_ports(), proto()); break;
      case GAR_OP_GA_CREATE_EG: {
        size_t li;
        if (op.a1 == GAR_NONE) {  // the listener made by the preceding GA_CREATE_LISTENER of this object
          auto it = made_listener.find({op.obj, op.a0});
          if (it == made_listener.end()) {
            st = Status::Err("GA_CREATE_EG without the listener its GA_CREATE_LISTENER should have made");
            break;
          }
          li = it->second;
        } else {
          li = rows.listener[op.a1].second;
        }
        size_t eg = SIZE_MAX;
        st = cloud.CreateEndpointGroup(op.a0, li, state.lbs[op.a2].arn, ip_preserve(), &eg);
        if (st) made_eg[{op.obj, op.a0}] = {li, eg};
        break;
      }
      case GAR_OP_GA_UPDATE_EG: {
        size_t acc = accel_of(op.a0), li, eg;
        if (op.a1 == GAR_PENDING) {
          if (op.a0 == GAR_PENDING) {
            li = made_chain[op.obj].listener;
            eg = made_chain[op.obj].eg;
          } else {
            auto it = made_eg.find({op.obj, op.a0});
            li = it == made_eg.end() ? SIZE_MAX : it->second.first;
            eg = it == made_eg.end() ? SIZE_MAX : it->second.second;
          }
          if (acc == SIZE_MAX || li == SIZE_MAX || eg == SIZE_MAX) {  // e.g. the Ingress chain whose listener error was swallowed
            st = Status::Err("endpoint group of an earlier op of this object does not exist");
            break;
          }
        } else {
          li = rows.eg[op.a1].second.first;
          eg = rows.eg[op.a1].second.second;
        }
        st = cloud.UpdateEndpointGroup(acc, li, eg, state.lbs[op.a2].arn, ip_preserve());
        break;
      }
      case GAR_OP_GA_DELETE_CHAIN:  // CleanupGlobalAccelerator (:254-272): endpoint group, listener, accelerator
        if (op.a2 != GAR_NONE) st = cloud.DeleteEndpointGroup(op.a0, rows.eg[op.a2].second.first, rows.eg[op.a2].second.second);
        if (st && op.a1 != GAR_NONE) st = cloud.DeleteListener(op.a0, rows.listener[op.a1].second);
        if (st) st = cloud.DeleteAccelerator(op.a0);
        break;
      case GAR_OP_R53_CREATE: {  // createMetadataRecordSet, then createRecordSet (route53.go:100-113)
        std::string host = Route53Hostname(*ob, op.sub & 0xFFFFFu);
        st = cloud.CreateMetadataRecordSet(op.a0, host, Route53OwnerValue(cluster, *ob));
        size_t rec = SIZE_MAX;
        if (st) st = cloud.CreateRecordSet(op.a0, host, state.accelerators[op.a1].dns, &rec);
        if (st) {
          made_record[{op.obj, host}] = rec;
          r53_created[{op.obj, op.sub >> 20}] = true;
        }
        break;
      }
      case GAR_OP_R53_UPSERT_A: {
        size_t rec;
        if (op.a2 == GAR_PENDING) {
          auto it = made_record.find({op.obj, Route53Hostname(*ob, op.sub & 0xFFFFFu)});
          if (it == made_record.end()) {
            st = Status::Err("alias record of an earlier op of this object does not exist");
            break;
          }
          rec = it->second;
        } else {
          rec = rows.record[op.a2].second;
        }
        st = cloud.UpdateRecordSet(op.a0, rec, state.accelerators[op.a1].dns);
        break;
      }
      case GAR_OP_R53_DELETE_RECORD: st = cloud.DeleteRecord(op.a0, rows.record[op.a1].second); break;
      default: continue;
    }
    if (st) rep.executed++;
    else fail(st);
  }
  // Events that hang off the status word, not off an op: a complete cleanup of an un-annotated object, and route53's
  // per-lbIngress "created" (only when that ensure returned no error)
  for (uint32_t i = 0; i < cs.n_objects; i++) {
    if ((filter & kExecGA) && (GAR_STATUS_EVENT(cs.status_ga[i]) & GAR_EV_DELETED) && !rep.Failed(GAR_CTRL_GA, i))
      rep.events.push_back({i, GAR_CTRL_GA, "Normal", "GlobalAcceleratorDeleted", "Global Accelerators are deleted"});
    if ((filter & kExecR53) && (GAR_STATUS_EVENT(cs.status_r53[i]) & GAR_EV_DELETED) && !rep.Failed(GAR_CTRL_R53, i))
      rep.events.push_back({i, GAR_CTRL_R53, "Normal", "Route53RecordDeleted", "Route53 record sets are deleted"});
  }
  for (auto &c : r53_created)
    if (!r53_j_failed.count(c.first)) rep.events.push_back({c.first.first, GAR_CTRL_R53, "Normal", "Route53RecourdCreated", "Route53 record set is created"});
  return rep;
}

// In-memory AWS.  Calls address existing resources by their position in the state that was packed; deletions are deferred to
// Commit() so positions stay valid while a change set is executed.  Fault injection: FailNth("CreateListener", 2) makes the
// second CreateListener call fail; FailAlways("UpdateRecordSet") every one.
class MockCloud : public Cloud {
 public:
  explicit MockCloud(CloudState *s) : s_(s) {}
  void FailNth(const std::string &call, int nth) { fail_nth_[call] = nth; }
  void FailAlways(const std::string &call) { fail_always_.insert(call); }
  std::vector<std::string> calls;  // every call, in order ("CreateAccelerator", ...), failed ones suffixed with "!"

  Status CreateAccelerator(const std::string &name, bool, const Tags &tags, size_t *accel) override {
    if (auto st = gate("CreateAccelerator"); !st) return st;
    Accelerator a;
    size_t id = ++serial_;
    a.arn = "arn:aws:globalaccelerator::1:accelerator/mock-" + std::to_string(id);
    a.name = name;
    a.dns = "mock" + std::to_string(id) + ".awsglobalaccelerator.com";
    a.tags = tags;
    s_->accelerators.push_back(a);
    *accel = s_->accelerators.size() - 1;
    return Status::OK();
  }
  Status UpdateAccelerator(size_t accel, const std::string &name, const Tags &tags) override {
    if (auto st = gate("UpdateAccelerator"); !st) return st;
    Accelerator &a = s_->accelerators[accel];
    a.enabled = true;
    a.name = name;
    // TagResource: replaces the value of existing keys, appends new ones (later duplicates of `tags` win)
    Tags fin;
    for (auto &t : tags) {
      bool seen = false;
      for (auto &f : fin)
        if (f.first == t.first) {
          f.second = t.second;
          seen = true;
        }
      if (!seen) fin.push_back(t);
    }
    Tags out;
    for (auto &t : a.tags) {
      bool replaced = false;
      for (auto &f : fin) replaced = replaced || f.first == t.first;
      if (!replaced) out.push_back(t);
    }
    for (auto &f : fin) out.push_back(f);
    a.tags = out;
    return Status::OK();
  }
  Status CreateListener(size_t accel, const std::vector<int32_t> &ports, uint8_t proto, size_t *listener) override {
    if (auto st = gate("CreateListener"); !st) return st;
    Accelerator &a = s_->accelerators[accel];
    Listener l;
    l.arn = a.arn + "/listener/n" + std::to_string(a.listeners.size());
    l.proto = proto;
    l.from_ports = ports;
    a.listeners.push_back(l);
    *listener = a.listeners.size() - 1;
    return Status::OK();
  }
  Status UpdateListener(size_t accel, size_t listener, const std::vector<int32_t> &ports, uint8_t proto) override {
    if (auto st = gate("UpdateListener"); !st) return st;
    Listener &l = s_->accelerators[accel].listeners[listener];
    l.from_ports = ports;
    l.proto = proto;
    return Status::OK();
  }
  Status CreateEndpointGroup(size_t accel, size_t listener, const std::string &lb_arn, bool, size_t *eg) override {
    if (auto st = gate("CreateEndpointGroup"); !st) return st;
    Listener &l = s_->accelerators[accel].listeners[listener];
    l.endpoint_groups.push_back(EndpointGroup{l.arn + "/endpoint-group/n" + std::to_string(l.endpoint_groups.size()), {lb_arn}});
    *eg = l.endpoint_groups.size() - 1;
    return Status::OK();
  }
  Status UpdateEndpointGroup(size_t accel, size_t listener, size_t eg, const std::string &lb_arn, bool) override {
    if (auto st = gate("UpdateEndpointGroup"); !st) return st;
    s_->accelerators[accel].listeners[listener].endpoint_groups[eg].endpoint_ids = {lb_arn};  // EndpointConfigurations replaces the list
    return Status::OK();
  }
  Status DeleteEndpointGroup(size_t accel, size_t listener, size_t eg) override {
    if (auto st = gate("DeleteEndpointGroup"); !st) return st;
    dead_eg_.insert({accel, {listener, eg}});
    return Status::OK();
  }
  Status DeleteListener(size_t accel, size_t listener) override {
    if (auto st = gate("DeleteListener"); !st) return st;
    dead_lis_.insert({accel, listener});
    return Status::OK();
  }
  Status DeleteAccelerator(size_t accel) override {
    if (auto st = gate("DeleteAccelerator"); !st) return st;
    dead_acc_.insert(accel);
    return Status::OK();
  }
  Status CreateMetadataRecordSet(size_t zone, const std::string &hostname, const std::string &owner_value) override {
    if (auto st = gate("CreateMetadataRecordSet"); !st) return st;
    RecordSet txt;
    txt.name = record_name(hostname);
    txt.type = GAR_RR_TXT;
    txt.values = {owner_value};
    s_->zones[zone].records.push_back(txt);
    return Status::OK();
  }
  Status CreateRecordSet(size_t zone, const std::string &hostname, const std::string &accel_dns, size_t *record) override {
    if (auto st = gate("CreateRecordSet"); !st) return st;
    RecordSet a;
    a.name = record_name(hostname);
    a.type = GAR_RR_A;
    a.has_alias = true;
    a.alias_dns = accel_dns + ".";
    s_->zones[zone].records.push_back(a);
    *record = s_->zones[zone].records.size() - 1;
    return Status::OK();
  }
  Status UpdateRecordSet(size_t zone, size_t record, const std::string &accel_dns) override {
    if (auto st = gate("UpdateRecordSet"); !st) return st;
    RecordSet &r = s_->zones[zone].records[record];
    r.type = GAR_RR_A;
    r.has_alias = true;
    r.alias_dns = accel_dns + ".";
    return Status::OK();
  }
  Status DeleteRecord(size_t zone, size_t record) override {
    if (auto st = gate("DeleteRecord"); !st) return st;
    dead_rec_.insert({zone, record});
    return Status::OK();
  }
  std::string AcceleratorArn(size_t accel) const override { return s_->accelerators[accel].arn; }
  // apply the deferred deletions
  void Commit() {
    for (auto it = dead_eg_.rbegin(); it != dead_eg_.rend(); ++it) {
      auto &egs = s_->accelerators[it->first].listeners[it->second.first].endpoint_groups;
      egs.erase(egs.begin() + (long)it->second.second);
    }
    for (auto it = dead_lis_.rbegin(); it != dead_lis_.rend(); ++it) s_->accelerators[it->first].listeners.erase(s_->accelerators[it->first].listeners.begin() + (long)it->second);
    for (auto it = dead_acc_.rbegin(); it != dead_acc_.rend(); ++it) s_->accelerators.erase(s_->accelerators.begin() + (long)*it);
    for (auto it = dead_rec_.rbegin(); it != dead_rec_.rend(); ++it) s_->zones[it->first].records.erase(s_->zones[it->first].records.begin() + (long)it->second);
    dead_eg_.clear();
    dead_lis_.clear();
    dead_acc_.clear();
    dead_rec_.clear();
  }

 private:
  static std::string record_name(const std::string &hostname) {  // Route53 stores '*' escaped and returns names with the trailing dot
    std::string name = hostname;
    size_t star = name.find('*');
    if (star != std::string::npos) name.replace(star, 1, "\\052");
    return name + ".";
  }
  Status gate(const std::string &call) {
    bool bad = fail_always_.count(call) != 0;
    auto it = fail_nth_.find(call);
    if (it != fail_nth_.end() && --it->second == 0) {
      bad = true;
      fail_nth_.erase(it);
    }
    calls.push_back(bad ? call + "!" : call);
    return bad ? Status::Err(call + ": injected failure") : Status::OK();
  }
  CloudState *s_;
  size_t serial_ = 0;
  std::map<std::string, int> fail_nth_;
  std::set<std::string> fail_always_;
  std::set<size_t> dead_acc_;
  std::set<std::pair<size_t, size_t>> dead_lis_;
  std::set<std::pair<size_t, std::pair<size_t, size_t>>> dead_eg_;
  std::set<std::pair<size_t, size_t>> dead_rec_;
};

}  // namespace garecon
