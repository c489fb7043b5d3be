// gar_rows.h — per-row decision logic of the diff engine (one call = one table row).
//
// Each function states the reference code it stands for (paths relative to the reference root).
// Everything here is GAR_HD: nvcc compiles it into the sm_100a kernels of libgarecon.so; tests/hostsim
// compiles the same text with g++ for CPU-side debugging of the device logic.
#pragma once

#include "gar_common.h"
#include "gar_json.h"

// ------------------------------------------------------------------ tables + work areas (device pointers)

struct DevTables {
  gar_objects o;
  gar_actual a;
  const u8 *cluster;  // --cluster-name, GAR_SLAB_PAD padded
  u32 cluster_len;
};

// internal per-object bits stored above the public GAR_DV_* byte in Work.derived (masked off before export)
enum {
  OBJ_HAS_NAME_ANN = 1u << 8,
  OBJ_HAS_TAGS_ANN = 1u << 9,
  OBJ_KEY_BAD = 1u << 10,  // ns/name are not laid out as "ns/name" (layout rule violated; engine reports GAR_E_INVALID)
};

// per-accelerator bits (from its tag list)
enum {
  ACC_MINE = 1u << 0,        // managed tag == "true" and cluster tag == --cluster-name (missing tag reads as "")
  ACC_OWNER_KEYED = 1u << 1, // owner tag is "service/<rest>" or "ingress/<rest>": can match an object key
  ACC_OWNER_INGRESS = 1u << 2,
  ACC_OWNER_3PART = 1u << 3, // <rest> contains exactly one '/': a key the controller could have written
  ACC_HAS_MANAGED = 1u << 4, ACC_HAS_OWNER = 1u << 5, ACC_HAS_THOST = 1u << 6,
};

// per-value-row class: is this ResourceRecord value an owner value of this cluster?
enum { VAL_NOT_OWNER = 0, VAL_OWNER_SERVICE = 1, VAL_OWNER_INGRESS = 2, VAL_OWNER_3PART = 4 };

struct Work {
  // objects
  u32 *derived;        // [n] GAR_DV_* | OBJ_*
  gar_str *ann_r53;    // [n] value of the route53-hostname annotation
  gar_str *ann_name;   // [n] value of global-accelerator-name
  gar_str *ann_tags;   // [n] value of global-accelerator-tags
  gar_str *ann_listen; // [n] value of alb.ingress.kubernetes.io/listen-ports
  u32 *dport_begin;    // [n+1] counts, then exclusive scan
  i32 *dports;
  // lbIngress rows
  u8 *tok_code;
  gar_str *tok_name, *tok_region;
  // accelerators
  u32 *acc_flags;
  gar_str *acc_owner_key;  // owner tag value minus the "service/" | "ingress/" prefix (refs into actual slab)
  gar_str *acc_owner;      // full owner tag value ("" if missing)
  gar_str *acc_thost;      // target-hostname tag value ("" if missing)
  gar_str *acc_managed;    // managed tag value
  // route53 expansion
  u32 *rec_zone;  // [n_records]
  u32 *val_rec;   // [n_values]
  u8 *val_cls;    // [n_values] VAL_*
  gar_str *val_key;  // [n_values] "<ns>/<name>" part of an owner value
  u8 *val_orphan;    // [n_values] 1 = owner value of this cluster whose object is not in the cache
  // indexes
  HashIdx ix_lb;     // (region, name) -> LB rows
  HashIdx ix_owner;  // (kind, "ns/name") -> accelerators that are ACC_MINE and ACC_OWNER_KEYED
  HashIdx ix_thost;  // target hostname -> accelerators that are ACC_MINE
  HashIdx ix_zone;   // zone name without the trailing dot -> zone rows
  HashIdx ix_val;    // (kind, "ns/name") -> owner value rows of this cluster
  HashIdx ix_alias;  // (zone, record name) -> alias record rows
  HashIdx ix_obj;    // (kind, "ns/name") -> object rows
  HashIdx ix_ovn;    // (zone, record name) -> orphan owner value rows
};

// ------------------------------------------------------------------ key hashes (build and probe sides must agree)

GAR_HD u64 key_hash_kinded(u32 kind, Str nsname) { return hmix(kind + 1, xxh64(nsname, 0)); }
GAR_HD u64 key_hash_lb(Str region, Str name) { return hmix(xxh64(region, 0), xxh64(name, 0)); }
GAR_HD u64 key_hash_zoned(u32 zone, Str name) { return hmix((u64)zone + 0x100, xxh64(name, 0)); }
GAR_HD u64 key_hash_str(Str s) { return xxh64(s, 0); }

// the workqueue key "ns/name" of an object row (cache.MetaNamespaceKeyFunc; reconcile.go:47).  Layout rule of
// gar_objects: name starts one byte after ns ends and that byte is '/'.
GAR_HD Str object_key(const DevTables &T, u32 i) {
  gar_str ns = T.o.obj_ns[i], nm = T.o.obj_name[i];
  return Str{T.o.slab + GAR_STR_OFF(ns), GAR_STR_LEN(ns) + 1 + GAR_STR_LEN(nm)};
}
GAR_HD bool object_key_ok(const DevTables &T, u32 i) {
  gar_str ns = T.o.obj_ns[i], nm = T.o.obj_name[i];
  u64 sep = GAR_STR_OFF(ns) + GAR_STR_LEN(ns);
  return GAR_STR_OFF(nm) == sep + 1 && sep < T.o.slab_len && T.o.slab[sep] == '/';
}

// ------------------------------------------------------------------ (a2) annotation-key filter / row classifier
//
// wasLoadBalancerService (globalaccelerator/service.go:18-26, route53/service.go:19-27), wasALBIngress
// (globalaccelerator/ingress.go:19-27), hasManagedAnnotation (controller.go:250-253), hasHostnameAnnotation
// (route53/controller.go:243-246), plus the annotation reads of global_accelerator.go:35-60,214,225,526 and
// listenerForService's protocol (global_accelerator.go:503-515).

#define ANN_PREFIX "aws-global-accelerator-controller.h3poteto.dev/"

GAR_HD bool proto_is(Str p, char a, char b, char c) {  // strings.ToLower(p) == "abc" for ASCII a, b, c
  return p.n == 3 && (p.p[0] | 0x20) == a && (p.p[1] | 0x20) == b && (p.p[2] | 0x20) == c;
}

GAR_HD void classify_object(const DevTables &T, const Work &W, u32 i) {
  const gar_objects &o = T.o;
  u32 kind = o.obj_kind[i];
  u32 dv = 0;
  bool has_lbtype = false, has_ingclass = false;
  gar_str r53 = 0, name = 0, tags = 0, listen = 0;
  bool has_listen = false;
  for (u32 k = o.obj_ann_begin[i]; k < o.obj_ann_begin[i + 1]; k++) {
    Str key = mkstr(o.slab, o.ann_key[k]);
    gar_str vref = o.ann_val[k];
    if (HAS_PREFIX_LIT(key, ANN_PREFIX)) {
      Str t = substr(key, sizeof(ANN_PREFIX) - 1, key.n - (u32)(sizeof(ANN_PREFIX) - 1));
      if (STREQ_LIT(t, "global-accelerator-managed")) {
        dv |= GAR_DV_GA_MANAGED;
      } else if (STREQ_LIT(t, "route53-hostname")) {
        dv |= GAR_DV_R53_ANNOTATED;
        r53 = vref;
      } else if (STREQ_LIT(t, "client-ip-preservation")) {
        if (STREQ_LIT(mkstr(o.slab, vref), "true")) dv |= GAR_DV_IP_PRESERVE;
        else dv &= ~(u32)GAR_DV_IP_PRESERVE;
      } else if (STREQ_LIT(t, "global-accelerator-name")) {
        dv |= OBJ_HAS_NAME_ANN;
        name = vref;
      } else if (STREQ_LIT(t, "global-accelerator-tags")) {
        dv |= OBJ_HAS_TAGS_ANN;
        tags = vref;
      } else if (STREQ_LIT(t, "ip-address-type")) {
        Str v = mkstr(o.slab, vref);
        if (STREQ_LIT(v, "ipv4") || STREQ_LIT(v, "IPV4")) dv |= GAR_DV_IPV4;
        else dv &= ~(u32)GAR_DV_IPV4;
      }
    } else if (STREQ_LIT(key, "service.beta.kubernetes.io/aws-load-balancer-type")) {
      has_lbtype = true;
    } else if (STREQ_LIT(key, "kubernetes.io/ingress.class")) {
      has_ingclass = true;
    } else if (STREQ_LIT(key, "alb.ingress.kubernetes.io/listen-ports")) {
      has_listen = true;
      listen = vref;
    }
  }
  if (kind == GAR_KIND_SERVICE) {
    bool lbsvc = o.obj_spec_type[i] == GAR_SVC_LOADBALANCER && (has_lbtype || (o.obj_flags[i] & GAR_OBJ_HAS_LB_CLASS));
    if (lbsvc) dv |= GAR_DV_GA_ELIGIBLE | GAR_DV_R53_ELIGIBLE;
    bool udp = false;  // listenerForService: the last port whose protocol is tcp/udp decides
    for (u32 p = o.obj_port_begin[i]; p < o.obj_port_begin[i + 1]; p++) {
      Str pr = mkstr(o.slab, o.port_proto[p]);
      if (proto_is(pr, 'u', 'd', 'p')) udp = true;
      else if (proto_is(pr, 't', 'c', 'p')) udp = false;
    }
    if (udp) dv |= GAR_DV_PROTO_UDP;
  } else {
    bool alb = has_ingclass;
    if ((o.obj_flags[i] & GAR_OBJ_HAS_INGRESS_CLASS) && STREQ_LIT(mkstr(o.slab, o.obj_ingress_class[i]), "alb")) alb = true;
    if (alb) dv |= GAR_DV_GA_ELIGIBLE;
    dv |= GAR_DV_R53_ELIGIBLE;  // the route53 controller does not filter Ingresses (route53/controller.go:130-148)
    if (has_listen) dv |= GAR_DV_PORTS_FROM_ANN;
  }
  if (!object_key_ok(T, i)) dv |= OBJ_KEY_BAD;
  W.derived[i] = dv;
  W.ann_r53[i] = r53;
  W.ann_name[i] = name;
  W.ann_tags[i] = tags;
  W.ann_listen[i] = listen;
}

// ------------------------------------------------------------------ (a3) LB hostname tokeniser
//
// DetectCloudProvider (pkg/cloudprovider/provider.go:8-17) followed by GetLBNameFromHostname and its helpers
// (pkg/cloudprovider/aws/load_balancer.go:32-93), as a byte scanner instead of six regexps.

GAR_HD bool is_word(u8 c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || c == '_'; }

// `^([\w\-]+)\-[\w]+$`: every byte is \w or '-', and the last '-' has >= 1 byte on both sides.
// Returns the length of group 1, or 0 when the regexp does not match.
GAR_HD u32 name_dash_id(Str s) {
  u32 last = GAR_NONE;
  for (u32 k = 0; k < s.n; k++) {
    u8 c = s.p[k];
    if (c == '-') last = k;
    else if (!is_word(c)) return 0;
  }
  if (last == GAR_NONE || last < 1 || last + 1 >= s.n) return 0;
  return last;
}

GAR_HD void tokenise_hostname(const DevTables &T, const Work &W, u32 row) {
  gar_str href = T.o.lbi_hostname[row];
  Str h = mkstr(T.o.slab, href);
  u64 base = GAR_STR_OFF(href);
  u8 code;
  gar_str name = 0, region = 0;
  // label boundaries: first three dots, and whether any dot exists
  u32 d1 = GAR_NONE, d2 = GAR_NONE, d3 = GAR_NONE, ndots = 0;
  for (u32 k = 0; k < h.n; k++)
    if (h.p[k] == '.') {
      if (ndots == 0) d1 = k;
      else if (ndots == 1) d2 = k;
      else if (ndots == 2) d3 = k;
      ndots++;
    }
  if (ndots == 0) {
    code = GAR_TOK_PANIC;  // parts[len(parts)-2] with a single part
  } else if (!(STREQ_LIT(h, "amazonaws.com") || HAS_SUFFIX_LIT(h, ".amazonaws.com"))) {
    code = GAR_TOK_NOT_AWS;
  } else if (HAS_SUFFIX_LIT(h, ".elb.amazonaws.com")) {
    // matchALBHostname: subdomain = label 0, region = label 1 (there are >= 4 labels here)
    Str sub = substr(h, 0, d1);
    region = GAR_STR(base + d1 + 1, d2 - d1 - 1);
    if (HAS_PREFIX_LIT(sub, "internal-")) {
      u32 g = name_dash_id(substr(sub, 9, sub.n - 9));
      if (g) {
        code = GAR_TOK_ALB_INTERNAL;
        name = GAR_STR(base + 9, g);
      } else {
        code = GAR_TOK_ERR_INTERNAL_ALB;
      }
    } else {
      u32 g = name_dash_id(sub);
      if (g) {
        code = GAR_TOK_ALB_PUBLIC;
        name = GAR_STR(base, g);
      } else {
        code = GAR_TOK_ERR_PUBLIC_ALB;
      }
    }
  } else {
    // nlbReg `\.elb\..+\.amazonaws\.com$`: some ".elb." at i with >= 1 byte between it and the final
    // ".amazonaws.com" (14 bytes), none of them '\n'.  The right-most candidate has the smallest middle.
    bool nlb = false;
    if (h.n >= 20) {
      for (u32 i = h.n - 20 + 1; i-- > 0;) {
        if (h.p[i] == '.' && h.p[i + 1] == 'e' && h.p[i + 2] == 'l' && h.p[i + 3] == 'b' && h.p[i + 4] == '.') {
          nlb = true;
          for (u32 k = i + 5; k < h.n - 14; k++)
            if (h.p[k] == '\n') nlb = false;
          break;
        }
      }
    }
    if (!nlb) {
      code = GAR_TOK_ERR_NOT_ELB;
    } else {
      // matchNLBHostname: subdomain = label 0, region = label 2 (>= 5 labels here)
      u32 g = name_dash_id(substr(h, 0, d1));
      if (g) {
        code = GAR_TOK_NLB;
        name = GAR_STR(base, g);
        region = GAR_STR(base + d2 + 1, d3 - d2 - 1);
      } else {
        code = GAR_TOK_ERR_NLB;
      }
    }
  }
  if (code > GAR_TOK_NLB) {
    name = 0;
    region = 0;
  }
  W.tok_code[row] = code;
  W.tok_name[row] = name;
  W.tok_region[row] = region;
}

// ------------------------------------------------------------------ (a5) accelerator tag digest
//
// tagsContainsAllValues builds map[key]value from the tag list (later duplicates win) and reads missing
// keys as "" (global_accelerator.go:559-570).  One pass per accelerator extracts the four system tags.

#define TAG_MANAGED "aws-global-accelerator-controller-managed"
#define TAG_OWNER "aws-global-accelerator-owner"
#define TAG_THOST "aws-global-accelerator-target-hostname"
#define TAG_CLUSTER "aws-global-accelerator-cluster"

GAR_HD u32 count_slashes(Str s) {
  u32 c = 0;
  for (u32 k = 0; k < s.n; k++) c += s.p[k] == '/';
  return c;
}

GAR_HD void digest_accelerator(const DevTables &T, const Work &W, u32 a) {
  const gar_actual &A = T.a;
  gar_str managed = 0, owner = 0, thost = 0, cluster = 0;
  u32 fl = 0;
  for (u32 t = A.acc_tag_begin[a]; t < A.acc_tag_begin[a + 1]; t++) {
    Str k = mkstr(A.slab, A.tag_key[t]);
    if (STREQ_LIT(k, TAG_MANAGED)) {
      managed = A.tag_val[t];
      fl |= ACC_HAS_MANAGED;
    } else if (STREQ_LIT(k, TAG_OWNER)) {
      owner = A.tag_val[t];
      fl |= ACC_HAS_OWNER;
    } else if (STREQ_LIT(k, TAG_THOST)) {
      thost = A.tag_val[t];
      fl |= ACC_HAS_THOST;
    } else if (STREQ_LIT(k, TAG_CLUSTER)) {
      cluster = A.tag_val[t];
    }
  }
  Str cl = mkstr(A.slab, cluster);
  if (STREQ_LIT(mkstr(A.slab, managed), "true") && streq(cl, Str{T.cluster, T.cluster_len})) fl |= ACC_MINE;
  Str ow = mkstr(A.slab, owner);
  gar_str key = 0;
  if (HAS_PREFIX_LIT(ow, "service/")) {
    fl |= ACC_OWNER_KEYED;
    key = GAR_STR(GAR_STR_OFF(owner) + 8, ow.n - 8);
  } else if (HAS_PREFIX_LIT(ow, "ingress/")) {
    fl |= ACC_OWNER_KEYED | ACC_OWNER_INGRESS;
    key = GAR_STR(GAR_STR_OFF(owner) + 8, ow.n - 8);
  }
  if ((fl & ACC_OWNER_KEYED) && count_slashes(mkstr(A.slab, key)) == 1) fl |= ACC_OWNER_3PART;
  W.acc_flags[a] = fl;
  W.acc_owner_key[a] = key;
  W.acc_owner[a] = owner;
  W.acc_thost[a] = thost;
  W.acc_managed[a] = managed;
}

// ------------------------------------------------------------------ (a9) owner values in Route53
//
// Route53OwnerValue (route53.go:18-20): "heritage=aws-global-accelerator-controller,cluster=<c>,<resource>/<ns>/<name>"
// including the double quotes.  A value can only ever equal an owner value computed for THIS cluster if it has
// that shape, so only such values enter the index, keyed like object keys.

#define R53_HERITAGE "\"heritage=aws-global-accelerator-controller,cluster="

GAR_HD void classify_value(const DevTables &T, const Work &W, u32 v) {
  const gar_actual &A = T.a;
  gar_str ref = A.val_value[v];
  Str s = mkstr(A.slab, ref);
  u8 cls = VAL_NOT_OWNER;
  gar_str key = 0;
  const u32 hl = (u32)(sizeof(R53_HERITAGE) - 1);
  u32 fixed = hl + T.cluster_len + 1 + 8 + 1;  // heritage + cluster + ',' + "service/" + closing quote
  if (s.n >= fixed && has_prefix_lit(s, R53_HERITAGE, hl) && streq(substr(s, hl, T.cluster_len), Str{T.cluster, T.cluster_len}) &&
      s.p[hl + T.cluster_len] == ',' && s.p[s.n - 1] == '"') {
    u32 ro = hl + T.cluster_len + 1;
    Str rest = substr(s, ro, s.n - 1 - ro);
    if (HAS_PREFIX_LIT(rest, "service/")) cls = VAL_OWNER_SERVICE;
    else if (HAS_PREFIX_LIT(rest, "ingress/")) cls = VAL_OWNER_INGRESS;
    if (cls) {
      key = GAR_STR(GAR_STR_OFF(ref) + ro + 8, rest.n - 8);
      if (count_slashes(mkstr(A.slab, key)) == 1) cls |= VAL_OWNER_3PART;
    }
  }
  W.val_cls[v] = cls;
  W.val_key[v] = key;
}

// ------------------------------------------------------------------ index probes

// GetLoadBalancer (load_balancer.go:13-30) on a client bound to `region` (aws.go:23-25): first row wins
GAR_HD u32 find_lb(const DevTables &T, const Work &W, Str region, Str name) {
  Cursor c = idx_open(W.ix_lb, key_hash_lb(region, name));
  for (u32 r; (r = idx_next(W.ix_lb, c)) != GAR_NONE;)
    if (streq(mkstr(T.a.slab, T.a.lb_name[r]), name) && streq(mkstr(T.a.slab, T.a.lb_region[r]), region)) return r;
  return GAR_NONE;
}

// ListGlobalAcceleratorByResource (global_accelerator.go:87-110) as a cursor over ix_owner
struct OwnerIter {
  Cursor c;
  Str key;
};
GAR_HD OwnerIter owner_open(const Work &W, u32 kind, Str key) { return OwnerIter{idx_open(W.ix_owner, key_hash_kinded(kind, key)), key}; }
GAR_HD u32 owner_next(const DevTables &T, const Work &W, u32 kind, OwnerIter &it) {
  for (u32 r; (r = idx_next(W.ix_owner, it.c)) != GAR_NONE;) {
    u32 fl = W.acc_flags[r];
    if ((((fl & ACC_OWNER_INGRESS) != 0) ? 1u : 0u) != kind) continue;
    if (streq(mkstr(T.a.slab, W.acc_owner_key[r]), it.key)) return r;
  }
  return GAR_NONE;
}

// ListGlobalAcceleratorByHostname (global_accelerator.go:62-85): number of matches (saturating at 2) and the first
GAR_HD u32 find_by_hostname(const DevTables &T, const Work &W, Str hostname, u32 *first) {
  Cursor c = idx_open(W.ix_thost, key_hash_str(hostname));
  u32 n = 0;
  *first = GAR_NONE;
  for (u32 r; (r = idx_next(W.ix_thost, c)) != GAR_NONE;) {
    if (!streq(mkstr(T.a.slab, W.acc_thost[r]), hostname)) continue;
    if (n == 0) *first = r;
    if (++n >= 2) break;
  }
  return n;
}

// GetHostedZone (route53.go:335-358) + parentDomain (:383-386)
GAR_HD u32 find_hosted_zone(const DevTables &T, const Work &W, Str hostname) {
  Str t = hostname;
  for (;;) {
    if (t.n == 0) return GAR_NONE;
    Cursor c = idx_open(W.ix_zone, key_hash_str(t));
    for (u32 z; (z = idx_next(W.ix_zone, c)) != GAR_NONE;) {
      Str zn = mkstr(T.a.slab, T.a.zone_name[z]);
      if (zn.n == t.n + 1 && streq(substr(zn, 0, t.n), t)) return z;  // zone.Name == target + "."
    }
    u32 k = 0;
    while (k < t.n && t.p[k] != '.') k++;
    if (k >= t.n) return GAR_NONE;  // single label: parent is ""
    t = substr(t, k + 1, t.n - k - 1);
  }
}

// replaceWildcards(name) == hostname + "."  (route53.go:360-371)
GAR_HD bool record_name_matches(Str name, Str hostname) {
  // first occurrence of \052 in name
  u32 w = GAR_NONE;
  for (u32 k = 0; k + 4 <= name.n; k++)
    if (name.p[k] == '\\' && name.p[k + 1] == '0' && name.p[k + 2] == '5' && name.p[k + 3] == '2') {
      w = k;
      break;
    }
  if (w == GAR_NONE) {
    return name.n == hostname.n + 1 && name.p[name.n - 1] == '.' && streq(substr(name, 0, hostname.n), hostname);
  }
  // unescaped = name[:w] + "*" + name[w+4:]; compare with hostname + "."
  if (name.n - 3 != hostname.n + 1) return false;
  if (name.p[name.n - 1] != '.') return false;  // note: name[w+4:] is never empty when this holds... checked below
  if (w >= hostname.n || hostname.p[w] != '*') return false;
  if (!streq(substr(name, 0, w), substr(hostname, 0, w))) return false;
  u32 tail = hostname.n - w - 1;  // bytes of hostname after '*'
  return streq(substr(name, w + 4, tail), substr(hostname, w + 1, tail));
}

// ------------------------------------------------------------------ op sink (count pass or write pass)

struct OpSink {
  gar_op *out;  // nullptr in the count pass
  u32 n;
  GAR_HD void put(u32 head, u32 obj, u32 sub, u32 a0, u32 a1, u32 a2) {
    if (out) {
      gar_op o;
      o.head = head;
      o.obj = obj;
      o.sub = sub;
      o.a0 = a0;
      o.a1 = a1;
      o.a2 = a2;
      out[n] = o;
    }
    n++;
  }
};

// ------------------------------------------------------------------ (a4)(a7) desired listener vs actual listener

// desired port list of an object: the input list, or the parsed annotation
struct PortList {
  const i32 *p;
  u32 n;
};
GAR_HD PortList desired_ports(const DevTables &T, const Work &W, u32 i) {
  if (W.derived[i] & GAR_DV_PORTS_FROM_ANN) return PortList{W.dports + W.dport_begin[i], W.dport_begin[i + 1] - W.dport_begin[i]};
  u32 b = T.o.obj_port_begin[i];
  return PortList{T.o.port_number + b, T.o.obj_port_begin[i + 1] - b};
}

// listenerPortChangedFrom{Service,Ingress} (global_accelerator.go:458-492): build count[port] over
// listener FromPorts and desired ports; changed iff some port has count <= 1, i.e. occurs exactly once in
// the concatenation of both lists.
GAR_HD bool ports_changed(PortList l, PortList d) {
  u32 n = l.n + d.n;
  for (u32 x = 0; x < n; x++) {
    i32 px = x < l.n ? l.p[x] : d.p[x - l.n];
    u32 cnt = 0;
    for (u32 y = 0; y < n && cnt < 2; y++) {
      i32 py = y < l.n ? l.p[y] : d.p[y - l.n];
      cnt += (py == px);
    }
    if (cnt <= 1) return true;
  }
  return false;
}

// acceleratorName(resource, obj) == *accelerator.Name  (global_accelerator.go:53-60, :417)
GAR_HD bool accel_name_matches(const DevTables &T, const Work &W, u32 i, u32 kind, Str acc_name) {
  if (W.derived[i] & OBJ_HAS_NAME_ANN) {
    Str n = mkstr(T.o.slab, W.ann_name[i]);
    if (n.n != 0) return streq(n, acc_name);
  }
  // resource + "-" + ns + "-" + name
  Str ns = mkstr(T.o.slab, T.o.obj_ns[i]), nm = mkstr(T.o.slab, T.o.obj_name[i]);
  if (acc_name.n != 7 + 1 + ns.n + 1 + nm.n) return false;
  const char *res = kind == GAR_KIND_SERVICE ? "service" : "ingress";
  for (u32 k = 0; k < 7; k++)
    if (acc_name.p[k] != (u8)res[k]) return false;
  if (acc_name.p[7] != '-' || acc_name.p[8 + ns.n] != '-') return false;
  return streq(substr(acc_name, 8, ns.n), ns) && streq(substr(acc_name, 9 + ns.n, nm.n), nm);
}

// value of tag `key` in the accelerator's tag list: later duplicates win, missing reads as ""
GAR_HD Str actual_tag(const DevTables &T, u32 acc, Str key) {
  Str v{T.a.slab, 0};
  for (u32 t = T.a.acc_tag_begin[acc]; t < T.a.acc_tag_begin[acc + 1]; t++)
    if (streq(mkstr(T.a.slab, T.a.tag_key[t]), key)) v = mkstr(T.a.slab, T.a.tag_val[t]);
  return v;
}

// One piece of the tags annotation: strings.Split(piece, "=") has exactly two parts
struct TagPiece {
  Str key, val;
  bool ok;
};
GAR_HD TagPiece tag_piece(Str piece) {
  u32 eq = GAR_NONE, neq = 0;
  for (u32 k = 0; k < piece.n; k++)
    if (piece.p[k] == '=') {
      if (neq == 0) eq = k;
      neq++;
    }
  TagPiece t;
  t.ok = neq == 1;
  if (t.ok) {
    t.key = substr(piece, 0, eq);
    t.val = substr(piece, eq + 1, piece.n - eq - 1);
  } else {
    t.key = t.val = Str{piece.p, 0};
  }
  return t;
}
// iterate the pieces of strings.Split(annotation, ",")
GAR_HD bool next_piece(Str all, u32 *pos, Str *piece) {
  if (*pos > all.n) return false;
  u32 b = *pos, k = b;
  while (k < all.n && all.p[k] != ',') k++;
  *piece = substr(all, b, k - b);
  *pos = k + 1;
  return true;
}

// acceleratorChanged (global_accelerator.go:412-437)
GAR_HD bool accelerator_changed(const DevTables &T, const Work &W, u32 i, u32 kind, u32 acc, Str lb_dns) {
  const gar_actual &A = T.a;
  if (!A.acc_enabled[acc]) return true;
  if (!accel_name_matches(T, W, i, kind, mkstr(A.slab, A.acc_name[acc]))) return true;
  // targetTags = {managed:"true", owner:resource/ns/name, target-hostname:lb dns} overlaid by the user tags of
  // the annotation (later pieces overwrite earlier ones and the three system keys).
  Str tags = (W.derived[i] & OBJ_HAS_TAGS_ANN) ? mkstr(T.o.slab, W.ann_tags[i]) : Str{T.o.slab, 0};
  bool user_managed = false, user_owner = false, user_thost = false;
  u32 pos = 0;
  Str piece;
  while (next_piece(tags, &pos, &piece)) {
    TagPiece tp = tag_piece(piece);
    if (!tp.ok) continue;
    if (STREQ_LIT(tp.key, TAG_MANAGED)) user_managed = true;
    else if (STREQ_LIT(tp.key, TAG_OWNER)) user_owner = true;
    else if (STREQ_LIT(tp.key, TAG_THOST)) user_thost = true;
    // overridden by a later piece with the same key?
    bool overridden = false;
    u32 pos2 = pos;
    Str piece2;
    while (next_piece(tags, &pos2, &piece2)) {
      TagPiece t2 = tag_piece(piece2);
      if (t2.ok && streq(t2.key, tp.key)) {
        overridden = true;
        break;
      }
    }
    if (overridden) continue;
    if (!streq(actual_tag(T, acc, tp.key), tp.val)) return true;
  }
  u32 fl = W.acc_flags[acc];
  if (!user_managed && !STREQ_LIT(mkstr(A.slab, W.acc_managed[acc]), "true")) return true;
  if (!user_owner) {
    // owner == resource/ns/name
    bool ok = (fl & ACC_OWNER_KEYED) && (((fl & ACC_OWNER_INGRESS) ? 1u : 0u) == kind) && streq(mkstr(A.slab, W.acc_owner_key[acc]), object_key(T, i));
    if (!ok) return true;
  }
  if (!user_thost && !streq(mkstr(A.slab, W.acc_thost[acc]), lb_dns)) return true;
  return false;
}

// CleanupGlobalAccelerator / listRelatedGlobalAccelerator (global_accelerator.go:254-288): the delete op of one accelerator
GAR_HD void put_delete_chain(const DevTables &T, OpSink &s, u32 obj, u32 kind, u32 acc) {
  const gar_actual &A = T.a;
  u32 lb = A.acc_lis_begin[acc], le = A.acc_lis_begin[acc + 1];
  u32 lis = GAR_NONE, eg = GAR_NONE;
  if (le - lb == 1) {
    lis = lb;
    u32 eb = A.lis_eg_begin[lis], ee = A.lis_eg_begin[lis + 1];
    if (ee - eb == 1) eg = eb;
  }
  s.put(GAR_OP_HEAD(GAR_OP_GA_DELETE_CHAIN, GAR_CTRL_GA, obj == GAR_NONE ? 0 : kind), obj, 0, acc, lis, eg);
}

// ------------------------------------------------------------------ (a6)(a7)(a8) Global Accelerator decisions of one object
//
// process{Service,Ingress}CreateOrUpdate (globalaccelerator/service.go:54-126, ingress.go:56-130) with
// EnsureGlobalAcceleratorFor* (global_accelerator.go:112-211) and updateGlobalAcceleratorFor* (:290-410) inlined.

GAR_HD u32 ga_reconcile(const DevTables &T, const Work &W, u32 i, OpSink &s) {
  const gar_objects &o = T.o;
  const gar_actual &A = T.a;
  u32 dv = W.derived[i];
  if (!(dv & GAR_DV_GA_ELIGIBLE)) return GAR_STATUS(GAR_ST_IGNORED, 0, 0);
  u32 jb = o.obj_lbi_begin[i], je = o.obj_lbi_begin[i + 1];
  if (je - jb < 1) return GAR_STATUS(GAR_ST_SKIP_NO_LB, 0, 0);
  u32 kind = o.obj_kind[i];
  Str okey = object_key(T, i);
  if (!(dv & GAR_DV_GA_MANAGED)) {
    OwnerIter it = owner_open(W, kind, okey);
    for (u32 acc; (acc = owner_next(T, W, kind, it)) != GAR_NONE;) put_delete_chain(T, s, i, kind, acc);
    return GAR_STATUS(GAR_ST_OK, 0, GAR_EV_DELETED);
  }
  u32 ev = 0;
  for (u32 j = 0; j < je - jb; j++) {
    u32 code = W.tok_code[jb + j];
    if (code == GAR_TOK_PANIC) return GAR_STATUS(GAR_ST_PANIC, 0, ev);
    if (code == GAR_TOK_NOT_AWS) continue;
    if (code >= GAR_TOK_ERR_NOT_ELB) return GAR_STATUS(GAR_ST_ERR_RETRY, GAR_D_NOT_ELB + (code - GAR_TOK_ERR_NOT_ELB), ev);
    Str hostname = mkstr(o.slab, o.lbi_hostname[jb + j]);
    u32 lb = find_lb(T, W, mkstr(o.slab, W.tok_region[jb + j]), mkstr(o.slab, W.tok_name[jb + j]));
    if (lb == GAR_NONE) return GAR_STATUS(GAR_ST_ERR_RETRY, GAR_D_LB_NOT_FOUND, ev);
    Str lb_dns = mkstr(A.slab, A.lb_dns[lb]);
    if (!streq(lb_dns, hostname)) return GAR_STATUS(GAR_ST_ERR_RETRY, GAR_D_LB_DNS_MISMATCH, ev);
    if (A.lb_state[lb] != GAR_LB_ACTIVE) return GAR_STATUS(GAR_ST_REQUEUE_30S, 0, ev);
    OwnerIter it = owner_open(W, kind, okey);
    u32 nacc = 0;
    for (u32 acc; (acc = owner_next(T, W, kind, it)) != GAR_NONE;) {
      nacc++;
      // updateGlobalAcceleratorFor{Service,Ingress}
      if (accelerator_changed(T, W, i, kind, acc, lb_dns)) s.put(GAR_OP_HEAD(GAR_OP_GA_UPDATE_ACCEL, GAR_CTRL_GA, kind), i, j, acc, lb, GAR_NONE);
      u32 lbeg = A.acc_lis_begin[acc], lend = A.acc_lis_begin[acc + 1];
      if (lend - lbeg > 1) return GAR_STATUS(GAR_ST_ERR_RETRY, GAR_D_TOO_MANY_LISTENERS, ev);
      if (lend == lbeg) {
        // the listener is created from the desired state, so neither change predicate fires on it; it has no
        // endpoint group yet, and the one created for it contains the LB (:298-345)
        s.put(GAR_OP_HEAD(GAR_OP_GA_CREATE_LISTENER, GAR_CTRL_GA, kind), i, j, acc, GAR_NONE, GAR_NONE);
        s.put(GAR_OP_HEAD(GAR_OP_GA_CREATE_EG, GAR_CTRL_GA, kind), i, j, acc, GAR_NONE, lb);
        continue;
      }
      u32 lis = lbeg;
      u32 want_proto = kind == GAR_KIND_SERVICE ? ((dv & GAR_DV_PROTO_UDP) ? GAR_PROTO_UDP : GAR_PROTO_TCP) : GAR_PROTO_TCP;
      bool changed = A.lis_proto[lis] != want_proto;  // :439-456
      if (!changed) {
        u32 pb = A.lis_pr_begin[lis];
        changed = ports_changed(PortList{A.pr_from + pb, A.lis_pr_begin[lis + 1] - pb}, desired_ports(T, W, i));
      }
      if (changed) s.put(GAR_OP_HEAD(GAR_OP_GA_UPDATE_LISTENER, GAR_CTRL_GA, kind), i, j, acc, lis, GAR_NONE);
      u32 eb = A.lis_eg_begin[lis], ee = A.lis_eg_begin[lis + 1];
      if (ee - eb > 1) return GAR_STATUS(GAR_ST_ERR_RETRY, GAR_D_TOO_MANY_EGS, ev);
      if (ee == eb) {
        s.put(GAR_OP_HEAD(GAR_OP_GA_CREATE_EG, GAR_CTRL_GA, kind), i, j, acc, lis, lb);
        continue;
      }
      u32 eg = eb;
      Str lb_arn = mkstr(A.slab, A.lb_arn[lb]);
      bool contains = false;  // endpointContainsLB (:494-501)
      for (u32 d = A.eg_ep_begin[eg]; d < A.eg_ep_begin[eg + 1] && !contains; d++) contains = streq(mkstr(A.slab, A.ep_id[d]), lb_arn);
      if (!contains) s.put(GAR_OP_HEAD(GAR_OP_GA_UPDATE_EG, GAR_CTRL_GA, kind), i, j, acc, eg, lb);
    }
    if (nacc == 0) {
      s.put(GAR_OP_HEAD(GAR_OP_GA_CREATE_CHAIN, GAR_CTRL_GA, kind), i, j, lb, GAR_NONE, GAR_NONE);
      ev |= GAR_EV_CREATED;
    }
  }
  return GAR_STATUS(GAR_ST_OK, 0, ev);
}

// ------------------------------------------------------------------ (a9)(a10) Route53 decisions of one object
//
// process{Service,Ingress}CreateOrUpdate of the route53 controller (route53/service.go:48-111, ingress.go:40-104),
// ensureRoute53 (route53.go:56-130) and CleanupRecordSet (:132-165).

// iterate the owner-value rows of one object key in ascending value-row order
struct ValIter {
  Cursor c;
  Str key;
  u32 kind;
};
GAR_HD ValIter val_open(const Work &W, u32 kind, Str key) { return ValIter{idx_open(W.ix_val, key_hash_kinded(kind, key)), key, kind}; }
GAR_HD u32 val_next(const DevTables &T, const Work &W, ValIter &it) {
  for (u32 v; (v = idx_next(W.ix_val, it.c)) != GAR_NONE;) {
    u32 cls = W.val_cls[v];
    if (((cls & VAL_OWNER_INGRESS) ? 1u : 0u) != it.kind) continue;
    if (streq(mkstr(T.a.slab, W.val_key[v]), it.key)) return v;
  }
  return GAR_NONE;
}

// first alias record with type A under (zone, name), or GAR_NONE  (rows of a bucket are ascending)
GAR_HD u32 first_alias_a(const DevTables &T, const Work &W, u32 zone, Str name) {
  Cursor c = idx_open(W.ix_alias, key_hash_zoned(zone, name));
  for (u32 r; (r = idx_next(W.ix_alias, c)) != GAR_NONE;) {
    if (W.rec_zone[r] != zone || T.a.rec_type[r] != GAR_RR_A) continue;
    if (streq(mkstr(T.a.slab, T.a.rec_name[r]), name)) return r;
  }
  return GAR_NONE;
}
// smallest alias record row > after under (zone, name), any type
GAR_HD u32 next_alias_any(const DevTables &T, const Work &W, u32 zone, Str name, u32 after /* GAR_NONE = none yet */) {
  Cursor c = idx_open(W.ix_alias, key_hash_zoned(zone, name));
  for (u32 r; (r = idx_next(W.ix_alias, c)) != GAR_NONE;) {
    if (after != GAR_NONE && r <= after) continue;
    if (W.rec_zone[r] != zone) continue;
    if (streq(mkstr(T.a.slab, T.a.rec_name[r]), name)) return r;
  }
  return GAR_NONE;
}

// CleanupRecordSet for one owner key: per zone, owned alias sets in record order, then owner metadata sets
GAR_HD void r53_cleanup(const DevTables &T, const Work &W, u32 obj, u32 kind, Str okey, OpSink &s) {
  const gar_actual &A = T.a;
  u32 head = GAR_OP_HEAD(GAR_OP_R53_DELETE_RECORD, GAR_CTRL_R53, obj == GAR_NONE ? 0 : kind);
  // value rows come out ascending, i.e. grouped by zone in zone order
  ValIter it = val_open(W, kind, okey);
  u32 v = val_next(T, W, it);
  while (v != GAR_NONE) {
    u32 zone = W.rec_zone[W.val_rec[v]];
    // phase 0: repeatedly take the smallest not-yet-emitted alias record whose name is one of the zone group's names
    u32 last = GAR_NONE;
    for (;;) {
      u32 best = GAR_NONE, best_v = GAR_NONE;
      ValIter g = val_open(W, kind, okey);
      for (u32 x; (x = val_next(T, W, g)) != GAR_NONE;) {
        u32 zx = W.rec_zone[W.val_rec[x]];
        if (zx < zone) continue;
        if (zx > zone) break;
        Str nm = mkstr(A.slab, A.rec_name[W.val_rec[x]]);
        u32 r = next_alias_any(T, W, zone, nm, last);
        // x ascending: the first value row that reaches a record is the one hostnameContains would hit first
        if (r != GAR_NONE && (best == GAR_NONE || r < best)) {
          best = r;
          best_v = x;
        }
      }
      if (best == GAR_NONE) break;
      s.put(head, obj, 0, zone, best, best_v);
      last = best;
    }
    // phase 1: one op per matching value of this zone
    u32 x = v;
    while (x != GAR_NONE && W.rec_zone[W.val_rec[x]] == zone) {
      s.put(head, obj, 1, zone, W.val_rec[x], x);
      x = val_next(T, W, it);
    }
    v = x;
  }
}

GAR_HD u32 r53_reconcile(const DevTables &T, const Work &W, u32 i, OpSink &s) {
  const gar_objects &o = T.o;
  const gar_actual &A = T.a;
  u32 dv = W.derived[i];
  if (!(dv & GAR_DV_R53_ELIGIBLE)) return GAR_STATUS(GAR_ST_IGNORED, 0, 0);
  u32 kind = o.obj_kind[i];
  Str okey = object_key(T, i);
  if (!(dv & GAR_DV_R53_ANNOTATED)) {
    r53_cleanup(T, W, i, kind, okey, s);
    return GAR_STATUS(GAR_ST_OK, 0, GAR_EV_DELETED);
  }
  Str hostnames = mkstr(o.slab, W.ann_r53[i]);
  u32 jb = o.obj_lbi_begin[i], je = o.obj_lbi_begin[i + 1];
  u32 ev = 0;
  for (u32 j = 0; j < je - jb; j++) {
    u32 code = W.tok_code[jb + j];
    if (code == GAR_TOK_PANIC) return GAR_STATUS(GAR_ST_PANIC, 0, ev);
    if (code == GAR_TOK_NOT_AWS) continue;
    if (code >= GAR_TOK_ERR_NOT_ELB) return GAR_STATUS(GAR_ST_ERR_RETRY, GAR_D_NOT_ELB + (code - GAR_TOK_ERR_NOT_ELB), ev);
    Str lb_hostname = mkstr(o.slab, o.lbi_hostname[jb + j]);
    u32 acc;
    u32 nacc = find_by_hostname(T, W, lb_hostname, &acc);
    if (nacc > 1) return GAR_STATUS(GAR_ST_REQUEUE_60S, GAR_D_ACCEL_MANY, ev);
    if (nacc == 0) return GAR_STATUS(GAR_ST_REQUEUE_60S, GAR_D_ACCEL_NONE, ev);
    Str acc_dns = mkstr(A.slab, A.acc_dns[acc]);
    bool created = false;
    u32 pos = 0, k = 0;
    Str hn;
    while (next_piece(hostnames, &pos, &hn)) {
      u32 zone = find_hosted_zone(T, W, hn);
      if (zone == GAR_NONE) return GAR_STATUS(GAR_ST_ERR_RETRY, GAR_D_NO_HOSTED_ZONE, ev);
      // findARecord over FindOwneredARecordSets: first (by record row) alias A record of the zone whose name is
      // owned by this object and unescapes to hostname + "."
      u32 rec = GAR_NONE;
      ValIter it = val_open(W, kind, okey);
      for (u32 v; (v = val_next(T, W, it)) != GAR_NONE;) {
        u32 vr = W.val_rec[v];
        u32 vz = W.rec_zone[vr];
        if (vz < zone) continue;
        if (vz > zone) break;
        Str nm = mkstr(A.slab, A.rec_name[vr]);
        if (!record_name_matches(nm, hn)) continue;
        u32 r = first_alias_a(T, W, zone, nm);
        if (r != GAR_NONE && (rec == GAR_NONE || r < rec)) rec = r;
      }
      if (rec == GAR_NONE) {
        s.put(GAR_OP_HEAD(GAR_OP_R53_CREATE, GAR_CTRL_R53, kind), i, GAR_R53_SUB(j, k), zone, acc, GAR_NONE);
        created = true;
      } else {
        // needRecordsUpdate (route53.go:373-381); rec is an alias record by construction
        Str al = mkstr(A.slab, A.rec_alias_dns[rec]);
        bool same = al.n == acc_dns.n + 1 && al.p[al.n - 1] == '.' && streq(substr(al, 0, acc_dns.n), acc_dns);
        if (!same) s.put(GAR_OP_HEAD(GAR_OP_R53_UPSERT_A, GAR_CTRL_R53, kind), i, GAR_R53_SUB(j, k), zone, acc, rec);
      }
      k++;
    }
    if (created) ev |= GAR_EV_CREATED;
  }
  return GAR_STATUS(GAR_ST_OK, 0, ev);
}

// ------------------------------------------------------------------ orphans (delete events of keys that left the cache)

GAR_HD bool object_in_cache(const DevTables &T, const Work &W, u32 kind, Str key) {
  Cursor c = idx_open(W.ix_obj, key_hash_kinded(kind, key));
  for (u32 r; (r = idx_next(W.ix_obj, c)) != GAR_NONE;)
    if (T.o.obj_kind[r] == kind && streq(object_key(T, r), key)) return true;
  return false;
}

// process{Service,Ingress}Delete of the globalaccelerator controller (service.go:28-52, ingress.go:29-54) for an
// accelerator whose owner key has no object: returns 1 and emits the delete op
GAR_HD u32 ga_orphan(const DevTables &T, const Work &W, u32 acc, OpSink &s) {
  u32 fl = W.acc_flags[acc];
  if (!(fl & ACC_MINE) || !(fl & ACC_OWNER_3PART)) return 0;
  u32 kind = (fl & ACC_OWNER_INGRESS) ? 1u : 0u;
  if (object_in_cache(T, W, kind, mkstr(T.a.slab, W.acc_owner_key[acc]))) return 0;
  put_delete_chain(T, s, GAR_NONE, 0, acc);
  return 1;
}

// is value row v an owner value of this cluster whose object left the cache?
GAR_HD void mark_orphan_value(const DevTables &T, const Work &W, u32 v) {
  u32 cls = W.val_cls[v];
  u8 orphan = 0;
  if (cls & VAL_OWNER_3PART) {
    u32 kind = (cls & VAL_OWNER_INGRESS) ? 1u : 0u;
    orphan = object_in_cache(T, W, kind, mkstr(T.a.slab, W.val_key[v])) ? 0 : 1;
  }
  W.val_orphan[v] = orphan;
}

// phase 0 of the orphan section for alias record r: one op per distinct orphan owner value that has a value
// row under the same (zone, name); a2 = the first such value row.  Rows of ix_ovn buckets are ascending.
GAR_HD void r53_orphan_alias(const DevTables &T, const Work &W, u32 r, OpSink &s) {
  const gar_actual &A = T.a;
  if (!A.rec_has_alias[r]) return;
  u32 zone = W.rec_zone[r];
  Str name = mkstr(A.slab, A.rec_name[r]);
  u64 h = key_hash_zoned(zone, name);
  Cursor c = idx_open(W.ix_ovn, h);
  for (u32 v; (v = idx_next(W.ix_ovn, c)) != GAR_NONE;) {
    u32 vr = W.val_rec[v];
    if (W.rec_zone[vr] != zone || !streq(mkstr(A.slab, A.rec_name[vr]), name)) continue;
    // skip if an earlier row of the bucket carries the same value under the same name
    Str val = mkstr(A.slab, A.val_value[v]);
    bool dup = false;
    Cursor c2 = idx_open(W.ix_ovn, h);
    for (u32 w; (w = idx_next(W.ix_ovn, c2)) != GAR_NONE && w < v;) {
      u32 wr = W.val_rec[w];
      if (W.rec_zone[wr] == zone && streq(mkstr(A.slab, A.rec_name[wr]), name) && streq(mkstr(A.slab, A.val_value[w]), val)) {
        dup = true;
        break;
      }
    }
    if (dup) continue;
    s.put(GAR_OP_HEAD(GAR_OP_R53_DELETE_RECORD, GAR_CTRL_R53, 0), GAR_NONE, 0, zone, r, v);
  }
}
