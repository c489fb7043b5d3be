// packer.hpp — host side of the C ABI, part 1: informer-cache objects + listed AWS resources
// -> the struct-of-arrays tables of include/garecon.h.
//
// In the reference these values are read one object at a time from the listers (globalaccelerator/controller.go:39-42)
// and from paginated AWS calls issued per reconcile (global_accelerator.go:624-652,789-813,885-907; load_balancer.go:13-30;
// route53.go:199-214,317-333).  A batch worker lists each of them ONCE and appends the rows here.  C++ because this image
// has no Go toolchain; the Go packer a maintainer would write has the same shape: append-only columns,
// one byte slab per table group, strings as (offset, len), CSR child ranges closed when the next parent starts.
//
// Layout rules the engine checks (gar_snapshot_load): "ns/name" is written once and obj_ns / obj_name point into it;
// CSR arrays are monotone and closed; every reference stays inside its slab.  Annotation keys and tag keys are interned
// (a handful of distinct strings referenced by every row).
//
// Strings are laid out COLUMN-MAJOR: every string column collects its bytes in its own buffer (rows in order) and Finish()
// concatenates the buffers into the table group's slab.  The engine accepts any layout (tests/test_hostsim_parity.py
// "slab layout"), but with column-major slabs consecutive threads of the row-local passes read consecutive bytes: the whole
// diff is 9 % faster than with strings row-major by parent.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/garecon.h"

namespace garecon {

// ---- a plain model of what the listers and the AWS list calls return (also the state of tests' mock cloud)
struct KObject {  // corev1.Service / networkingv1.Ingress, the fields the path reads
  uint8_t kind = GAR_KIND_SERVICE;
  std::string ns, name;
  uint8_t spec_type = GAR_SVC_LOADBALANCER;  // Service only
  bool has_lb_class = false;                 // Service: spec.loadBalancerClass != nil
  bool has_ingress_class = false;            // Ingress: spec.ingressClassName != nil
  std::string ingress_class;
  std::vector<std::pair<std::string, std::string>> annotations;
  std::vector<std::string> lb_hostnames;               // status.loadBalancer.ingress[].hostname
  std::vector<std::pair<int32_t, std::string>> ports;  // Service: spec.ports (number, protocol); Ingress: backend port numbers
};
struct EndpointGroup {
  std::string arn;
  std::vector<std::string> endpoint_ids;
};
struct Listener {
  std::string arn;
  uint8_t proto = GAR_PROTO_TCP;
  std::vector<int32_t> from_ports;
  std::vector<EndpointGroup> endpoint_groups;
};
struct Accelerator {
  std::string arn, name, dns;
  bool enabled = true;
  std::vector<std::pair<std::string, std::string>> tags;
  std::vector<Listener> listeners;
};
struct LoadBalancer {
  std::string region, name, dns, arn;
  uint8_t state = GAR_LB_ACTIVE;
};
struct RecordSet {
  std::string name;  // as AWS returns it: trailing dot, '*' as \052
  uint8_t type = GAR_RR_OTHER;
  bool has_alias = false;
  std::string alias_dns;
  std::vector<std::string> values;
};
struct HostedZone {
  std::string id, name;
  std::vector<RecordSet> records;
};
struct CloudState {  // everything the path lists from AWS
  std::vector<LoadBalancer> lbs;
  std::vector<Accelerator> accelerators;
  std::vector<HostedZone> zones;
};

class Packer {
 public:
  Packer() { Reset(); }
  void Reset() {
    *this = Packer(0);
    ann_b_.push_back(0); lbi_b_.push_back(0); port_b_.push_back(0);
    tag_b_.push_back(0); lis_b_.push_back(0); pr_b_.push_back(0); eg_b_.push_back(0); ep_b_.push_back(0); rec_b_.push_back(0); val_b_.push_back(0);
  }

  // ---- desired side: one call per object, in lister order
  uint32_t AddObject(const KObject &k) {
    obj_kind_.push_back(k.kind);
    obj_spec_.push_back(k.kind == GAR_KIND_SERVICE ? k.spec_type : 0);
    obj_flags_.push_back((k.has_lb_class ? GAR_OBJ_HAS_LB_CLASS : 0) | (k.has_ingress_class ? GAR_OBJ_HAS_INGRESS_CLASS : 0));
    gar_str key = put(OKEY, k.ns + "/" + k.name);  // cache.MetaNamespaceKeyFunc (reconcile.go:47): one string, two views
    obj_ns_.push_back(GAR_STR(GAR_STR_OFF(key), k.ns.size()));
    obj_name_.push_back(GAR_STR(GAR_STR_OFF(key) + k.ns.size() + 1, k.name.size()));
    obj_icls_.push_back(k.has_ingress_class ? put(OICLS, k.ingress_class) : 0);
    for (auto &a : k.annotations) {
      ann_key_.push_back(intern(OKEYS, okeys_, a.first));
      ann_val_.push_back(put(OANNV, a.second));
    }
    ann_b_.push_back((uint32_t)ann_key_.size());
    for (auto &h : k.lb_hostnames) lbi_host_.push_back(put(OHOST, h));
    lbi_b_.push_back((uint32_t)lbi_host_.size());
    for (auto &p : k.ports) {
      port_num_.push_back(p.first);
      port_proto_.push_back(p.second.empty() ? 0 : intern(OKEYS, okeys_, p.second));
    }
    port_b_.push_back((uint32_t)port_num_.size());
    return (uint32_t)obj_kind_.size() - 1;
  }
  // ---- actual side
  uint32_t AddLoadBalancer(const LoadBalancer &lb) {
    lb_region_.push_back(intern(AKEYS, akeys_, lb.region));
    lb_name_.push_back(put(ALBNAME, lb.name));
    lb_dns_.push_back(put(ALBDNS, lb.dns));
    lb_arn_.push_back(put(ALBARN, lb.arn));
    lb_state_.push_back(lb.state);
    return (uint32_t)lb_state_.size() - 1;
  }
  uint32_t AddAccelerator(const Accelerator &a) {
    acc_name_.push_back(put(AACCNAME, a.name));
    acc_dns_.push_back(put(AACCDNS, a.dns));
    acc_enabled_.push_back(a.enabled ? 1 : 0);
    for (auto &t : a.tags) {
      tag_key_.push_back(intern(AKEYS, akeys_, t.first));
      tag_val_.push_back(put(ATAGV, t.second));
    }
    tag_b_.push_back((uint32_t)tag_key_.size());
    for (auto &l : a.listeners) {
      lis_proto_.push_back(l.proto);
      for (int32_t p : l.from_ports) pr_from_.push_back(p);
      pr_b_.push_back((uint32_t)pr_from_.size());
      for (auto &g : l.endpoint_groups) {
        for (auto &id : g.endpoint_ids) ep_id_.push_back(put(AEP, id));
        ep_b_.push_back((uint32_t)ep_id_.size());
      }
      eg_b_.push_back((uint32_t)ep_b_.size() - 1);
    }
    lis_b_.push_back((uint32_t)lis_proto_.size());
    return (uint32_t)acc_enabled_.size() - 1;
  }
  uint32_t AddZone(const HostedZone &z) {
    close_zone();
    open_zone(z);
    for (auto &r : z.records) add_record(r);
    close_zone();
    return (uint32_t)zone_name_.size() - 1;
  }
  void AddCloud(const CloudState &c) {
    for (auto &lb : c.lbs) AddLoadBalancer(lb);
    for (auto &a : c.accelerators) AddAccelerator(a);
    for (auto &z : c.zones) AddZone(z);
  }

  // ---- paginated ingestion: one call per page of the reference's pagers, in the order the pages arrive.  Row order IS list
  // order (the change set's canonical order and "first row wins" lookups depend on it), so pages are append-only:
  //   DescribeLoadBalancers pages, every region listed                  load_balancer.go:13-30
  //   ListAccelerators pages (100 per page), each accelerator with its ListTagsForResource / ListListeners /
  //   ListEndpointGroups results attached                               global_accelerator.go:624-652,789-813,885-907
  //   ListHostedZones pages (100), then per zone its ListResourceRecordSets pages (300)   route53.go:199-214,317-333
  // Returns the row of the page's first element.
  uint32_t AddLoadBalancerPage(const std::vector<LoadBalancer> &page) {
    uint32_t first = (uint32_t)lb_state_.size();
    for (auto &lb : page) AddLoadBalancer(lb);
    return first;
  }
  uint32_t AddAcceleratorPage(const std::vector<Accelerator> &page) {
    uint32_t first = (uint32_t)acc_enabled_.size();
    for (auto &a : page) AddAccelerator(a);
    return first;
  }
  // zone rows are fixed by the ListHostedZones pages; record pages follow zone by zone (AddRecordSetPage)
  uint32_t AddHostedZonePage(const std::vector<HostedZone> &page) {
    uint32_t first = (uint32_t)pending_zones_.size();
    for (auto &z : page) pending_zones_.push_back(HostedZone{z.id, z.name, {}});
    return first;
  }
  // One ListResourceRecordSets page of zone row `zone`.  Zones must be visited in ascending row order (records are stored
  // zone-major); a zone's pages are consecutive; zones that are skipped have no record sets.  false = out of order.
  bool AddRecordSetPage(uint32_t zone, const std::vector<RecordSet> &page) {
    if (zone >= pending_zones_.size() || zone < zone_name_.size() - (zone_open_ ? 1 : 0)) return false;
    if (!(zone_open_ && zone == zone_name_.size() - 1)) {
      close_zone();
      while (zone_name_.size() < zone) open_zone(pending_zones_[zone_name_.size()]), close_zone();  // zones without records
      open_zone(pending_zones_[zone]);
    }
    for (auto &r : page) add_record(r);
    return true;
  }

  // ---- publish: the two table structs point into this object's buffers (valid until the next Add* / Reset)
  void Finish() {
    close_zone();
    while (zone_name_.size() < pending_zones_.size()) open_zone(pending_zones_[zone_name_.size()]), close_zone();  // listed zones without record pages
    // concatenate the column buffers into the two slabs and rebase every reference (no padding needed on the host side:
    // gar_snapshot_load copies and pads on the device)
    uint64_t base[NCOLS];
    os_.clear();
    as_.clear();
    for (int c = 0; c < NCOLS; c++) {
      std::string &dst = c < AFIRST ? os_ : as_;
      base[c] = dst.size();
      dst += col_[c];
    }
    auto rebase = [&](std::vector<gar_str> &v, int c) {
      for (auto &r : v) r += base[c];  // the offset lives in the low bits (empty strings move too: obj_ns / obj_name must stay adjacent)
    };
    rebase(obj_ns_, OKEY); rebase(obj_name_, OKEY); rebase(obj_icls_, OICLS); rebase(ann_key_, OKEYS); rebase(ann_val_, OANNV);
    rebase(lbi_host_, OHOST); rebase(port_proto_, OKEYS);
    rebase(lb_region_, AKEYS); rebase(lb_name_, ALBNAME); rebase(lb_dns_, ALBDNS); rebase(lb_arn_, ALBARN);
    rebase(acc_name_, AACCNAME); rebase(acc_dns_, AACCDNS); rebase(tag_key_, AKEYS); rebase(tag_val_, ATAGV); rebase(ep_id_, AEP);
    rebase(zone_name_, AZONE); rebase(rec_name_, ARECNAME); rebase(rec_alias_, AALIAS); rebase(val_value_, AVAL);
    for (auto &c : col_) std::string().swap(c);
    os_len_ = os_.size();
    as_len_ = as_.size();
    o_ = gar_objects{};
    o_.n_objects = (uint32_t)obj_kind_.size();
    o_.obj_kind = obj_kind_.data(); o_.obj_spec_type = obj_spec_.data(); o_.obj_flags = obj_flags_.data();
    o_.obj_ns = obj_ns_.data(); o_.obj_name = obj_name_.data(); o_.obj_ingress_class = obj_icls_.data();
    o_.obj_ann_begin = ann_b_.data(); o_.obj_lbi_begin = lbi_b_.data(); o_.obj_port_begin = port_b_.data();
    o_.n_ann = (uint32_t)ann_key_.size(); o_.ann_key = ann_key_.data(); o_.ann_val = ann_val_.data();
    o_.n_lbi = (uint32_t)lbi_host_.size(); o_.lbi_hostname = lbi_host_.data();
    o_.n_ports = (uint32_t)port_num_.size(); o_.port_number = port_num_.data(); o_.port_proto = port_proto_.data();
    o_.slab = (const uint8_t *)os_.data(); o_.slab_len = os_len_;
    a_ = gar_actual{};
    a_.n_lbs = (uint32_t)lb_state_.size();
    a_.lb_region = lb_region_.data(); a_.lb_name = lb_name_.data(); a_.lb_dns = lb_dns_.data(); a_.lb_arn = lb_arn_.data(); a_.lb_state = lb_state_.data();
    a_.n_accels = (uint32_t)acc_enabled_.size();
    a_.acc_name = acc_name_.data(); a_.acc_dns = acc_dns_.data(); a_.acc_enabled = acc_enabled_.data();
    a_.acc_tag_begin = tag_b_.data(); a_.acc_lis_begin = lis_b_.data();
    a_.n_tags = (uint32_t)tag_key_.size(); a_.tag_key = tag_key_.data(); a_.tag_val = tag_val_.data();
    a_.n_listeners = (uint32_t)lis_proto_.size(); a_.lis_proto = lis_proto_.data(); a_.lis_pr_begin = pr_b_.data(); a_.lis_eg_begin = eg_b_.data();
    a_.n_port_ranges = (uint32_t)pr_from_.size(); a_.pr_from = pr_from_.data();
    a_.n_egs = (uint32_t)ep_b_.size() - 1; a_.eg_ep_begin = ep_b_.data();
    a_.n_endpoints = (uint32_t)ep_id_.size(); a_.ep_id = ep_id_.data();
    a_.n_zones = (uint32_t)zone_name_.size(); a_.zone_name = zone_name_.data(); a_.zone_rec_begin = rec_b_.data();
    a_.n_records = (uint32_t)rec_name_.size();
    a_.rec_name = rec_name_.data(); a_.rec_type = rec_type_.data(); a_.rec_has_alias = rec_has_alias_.data(); a_.rec_alias_dns = rec_alias_.data();
    a_.rec_val_begin = val_b_.data();
    a_.n_values = (uint32_t)val_value_.size(); a_.val_value = val_value_.data();
    a_.slab = (const uint8_t *)as_.data(); a_.slab_len = as_len_;
  }
  const gar_objects *objects() const { return &o_; }
  const gar_actual *actual() const { return &a_; }
  uint64_t bytes() const { return os_len_ + as_len_ + 8 * (obj_ns_.size() * 3 + ann_key_.size() * 2 + lbi_host_.size() + tag_key_.size() * 2 + rec_name_.size() * 2); }

 private:
  explicit Packer(int) {}
  using KeyMap = std::unordered_map<std::string, gar_str>;
  // string columns: objects side first, then the actual side (AFIRST); *KEYS hold the interned strings
  enum Col { OKEYS, OKEY, OICLS, OANNV, OHOST, AFIRST, AKEYS = AFIRST, ALBNAME, ALBDNS, ALBARN, AACCNAME, AACCDNS, ATAGV, AEP, AZONE, ARECNAME, AALIAS, AVAL, NCOLS };
  gar_str put(int c, std::string_view s) {  // column-local offset until Finish()
    gar_str r = GAR_STR(col_[c].size(), s.size());
    col_[c].append(s.data(), s.size());
    return r;
  }
  gar_str intern(int c, KeyMap &m, const std::string &s) {
    auto it = m.find(s);
    if (it != m.end()) return it->second;
    gar_str r = put(c, s);
    if (m.size() < 4096) m.emplace(s, r);
    return r;
  }
  void open_zone(const HostedZone &z) {
    zone_name_.push_back(put(AZONE, z.name));
    zone_open_ = true;
  }
  void close_zone() {
    if (!zone_open_) return;
    rec_b_.push_back((uint32_t)rec_name_.size());
    zone_open_ = false;
  }
  void add_record(const RecordSet &r) {
    rec_name_.push_back(put(ARECNAME, r.name));
    rec_type_.push_back(r.type);
    rec_has_alias_.push_back(r.has_alias ? 1 : 0);
    rec_alias_.push_back(r.has_alias ? put(AALIAS, r.alias_dns) : 0);
    for (auto &v : r.values) val_value_.push_back(put(AVAL, v));
    val_b_.push_back((uint32_t)val_value_.size());
  }
  bool zone_open_ = false;
  std::vector<HostedZone> pending_zones_;  // zones listed by AddHostedZonePage (metadata only)
  std::string col_[NCOLS];
  std::string os_, as_;
  uint64_t os_len_ = 0, as_len_ = 0;
  KeyMap okeys_, akeys_;
  std::vector<uint8_t> obj_kind_, obj_spec_, obj_flags_, lb_state_, acc_enabled_, lis_proto_, rec_type_, rec_has_alias_;
  std::vector<gar_str> obj_ns_, obj_name_, obj_icls_, ann_key_, ann_val_, lbi_host_, port_proto_, lb_region_, lb_name_, lb_dns_, lb_arn_, acc_name_, acc_dns_,
      tag_key_, tag_val_, ep_id_, zone_name_, rec_name_, rec_alias_, val_value_;
  std::vector<uint32_t> ann_b_, lbi_b_, port_b_, tag_b_, lis_b_, pr_b_, eg_b_, ep_b_, rec_b_, val_b_;
  std::vector<int32_t> port_num_, pr_from_;
  gar_objects o_{};
  gar_actual a_{};
};

}  // namespace garecon
