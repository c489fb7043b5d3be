// gar_pipeline.h — the diff as a sequence of data-parallel stages over a Backend.
//
// Backend = how a stage runs: the CUDA backend (gar_engine.cu) launches sm_100a kernels on a stream; the
// host-simulation backend (tests/hostsim) runs the same functors in loops so the CPU-only test tier can
// exercise the device logic.  Stages, functors and their order live here once.
//
// Backend interface:
//   template <class F> void for_each(const char *name, u32 n, const F &f);  // f(i) for i in [0, n)
//   void exclusive_scan(u32 *data, u32 n);                                  // in place
//   void sort_pairs(u32 *keys, u32 *vals, u32 *keys_alt, u32 *vals_alt, u32 n, int bits);  // stable, result in keys/vals
//   void fill32(u32 *p, u32 value, size_t n);
//   void *ensure(int slot, size_t bytes);                                   // scratch buffer `slot`, at least `bytes`
//   void download(void *host_dst, const void *dev_src, size_t bytes);       // blocking
#pragma once

#include "gar_rows.h"

#if defined(__CUDA_ARCH__)
#define GAR_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define GAR_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

// ------------------------------------------------------------------ scratch slots

enum Slot {
  S_DERIVED, S_OFLAGS, S_ANN_R53, S_ANN_NAME, S_ANN_TAGS, S_ANN_LISTEN, S_DPORT_BEGIN, S_DPORTS,
  S_TOK_CODE, S_TOK_NAME, S_TOK_REGION,
  S_ACC_FLAGS, S_ACC_OWNER_KEY, S_ACC_OWNER, S_ACC_THOST, S_ACC_MANAGED,
  S_REC_ZONE, S_VAL_REC, S_VAL_CLS, S_VAL_KEY, S_VAL_ORPHAN,
  S_IX_LB, S_IX_OWNER = S_IX_LB + 3, S_IX_THOST = S_IX_OWNER + 3, S_IX_ZONE = S_IX_THOST + 3, S_IX_VAL = S_IX_ZONE + 3,
  S_IX_ALIAS = S_IX_VAL + 3, S_IX_OBJ = S_IX_ALIAS + 3, S_IX_OVN = S_IX_OBJ + 3,
  S_SORT_KEYS = S_IX_OVN + 3, S_SORT_VALS, S_SORT_KEYS_ALT, S_SORT_VALS_ALT, S_SORT_TAGS,
  S_COUNTS, S_STATUS_GA, S_STATUS_R53, S_OPS, S_ERRFLAG,
  S_NSLOTS
};

// ------------------------------------------------------------------ functors (one per stage)

struct FClassify {
  DevTables T;
  Work W;
  u32 *derived_public;
  u8 *oflags;
  u32 *errflag;
  GAR_HD void operator()(u32 i) const {
    classify_object(T, W, i);
    u32 dv = W.derived[i];
    if (dv & OBJ_KEY_BAD) GAR_ATOMIC_ADD(errflag, 1u);
    derived_public[i] = dv & 0xFFu;
  }
};
struct FTokenise {
  DevTables T;
  Work W;
  GAR_HD void operator()(u32 i) const { tokenise_hostname(T, W, i); }
};
struct FDigestAccel {
  DevTables T;
  Work W;
  GAR_HD void operator()(u32 i) const { digest_accelerator(T, W, i); }
};
// child row -> parent row for a CSR (binary search on the begin array)
struct FExpand {
  const u32 *begin;
  u32 nparents;
  u32 *parent_of;
  GAR_HD void operator()(u32 c) const {
    u32 lo = 0, hi = nparents;  // last p with begin[p] <= c
    while (hi - lo > 1) {
      u32 mid = (lo + hi) >> 1;
      if (begin[mid] <= c) lo = mid;
      else hi = mid;
    }
    parent_of[c] = lo;
  }
};
struct FClassifyValue {
  DevTables T;
  Work W;
  GAR_HD void operator()(u32 v) const { classify_value(T, W, v); }
};
struct FJsonCount {
  DevTables T;
  Work W;
  GAR_HD void operator()(u32 i) const {
    u32 n = 0;
    if (W.derived[i] & GAR_DV_PORTS_FROM_ANN) {
      u32 stk[GAR_JSON_STACK_WORDS];
      int c = json_listen_ports(mkstr(T.o.slab, W.ann_listen[i]), nullptr, stk);
      n = c < 0 ? 0u : (u32)c;
    }
    W.dport_begin[i] = n;
  }
};
struct FJsonWrite {
  DevTables T;
  Work W;
  GAR_HD void operator()(u32 i) const {
    if (!(W.derived[i] & GAR_DV_PORTS_FROM_ANN)) return;
    if (W.dport_begin[i + 1] == W.dport_begin[i]) return;
    u32 stk[GAR_JSON_STACK_WORDS];
    json_listen_ports(mkstr(T.o.slab, W.ann_listen[i]), W.dports + W.dport_begin[i], stk);
  }
};

// --- index key functors: write bucket key (or the sentinel `nb` for rows that are not indexed) and the tag
struct IdxOut {
  u32 *keys, *vals, *tags;
  u32 mask, nb;
  GAR_HD void put(u32 i, bool valid, u64 h) const {
    keys[i] = valid ? hash_bucket(h, mask) : nb;
    vals[i] = i;
    tags[i] = hash_tag(h);
  }
};
struct FKeyLb {
  DevTables T;
  IdxOut o;
  GAR_HD void operator()(u32 i) const { o.put(i, true, key_hash_lb(mkstr(T.a.slab, T.a.lb_region[i]), mkstr(T.a.slab, T.a.lb_name[i]))); }
};
struct FKeyOwner {
  DevTables T;
  Work W;
  IdxOut o;
  GAR_HD void operator()(u32 i) const {
    u32 fl = W.acc_flags[i];
    bool valid = (fl & ACC_MINE) && (fl & ACC_OWNER_KEYED);
    u64 h = valid ? key_hash_kinded((fl & ACC_OWNER_INGRESS) ? 1u : 0u, mkstr(T.a.slab, W.acc_owner_key[i])) : 0;
    o.put(i, valid, h);
  }
};
struct FKeyThost {
  DevTables T;
  Work W;
  IdxOut o;
  GAR_HD void operator()(u32 i) const {
    bool valid = (W.acc_flags[i] & ACC_MINE) != 0;
    o.put(i, valid, valid ? key_hash_str(mkstr(T.a.slab, W.acc_thost[i])) : 0);
  }
};
struct FKeyZone {
  DevTables T;
  IdxOut o;
  GAR_HD void operator()(u32 i) const {
    Str zn = mkstr(T.a.slab, T.a.zone_name[i]);
    bool valid = zn.n >= 1 && zn.p[zn.n - 1] == '.';
    o.put(i, valid, valid ? key_hash_str(substr(zn, 0, zn.n - 1)) : 0);
  }
};
struct FKeyVal {
  DevTables T;
  Work W;
  IdxOut o;
  GAR_HD void operator()(u32 v) const {
    u32 cls = W.val_cls[v];
    bool valid = cls != VAL_NOT_OWNER;
    o.put(v, valid, valid ? key_hash_kinded((cls & VAL_OWNER_INGRESS) ? 1u : 0u, mkstr(T.a.slab, W.val_key[v])) : 0);
  }
};
struct FKeyAlias {
  DevTables T;
  Work W;
  IdxOut o;
  GAR_HD void operator()(u32 r) const {
    bool valid = T.a.rec_has_alias[r] != 0;
    o.put(r, valid, valid ? key_hash_zoned(W.rec_zone[r], mkstr(T.a.slab, T.a.rec_name[r])) : 0);
  }
};
struct FKeyObj {
  DevTables T;
  Work W;
  IdxOut o;
  GAR_HD void operator()(u32 i) const {
    bool valid = !(W.derived[i] & OBJ_KEY_BAD);
    o.put(i, valid, valid ? key_hash_kinded(T.o.obj_kind[i], object_key(T, i)) : 0);
  }
};
struct FKeyOvn {
  DevTables T;
  Work W;
  IdxOut o;
  GAR_HD void operator()(u32 v) const {
    bool valid = W.val_orphan[v] != 0;
    u32 r = W.val_rec[v];
    o.put(v, valid, valid ? key_hash_zoned(W.rec_zone[r], mkstr(T.a.slab, T.a.rec_name[r])) : 0);
  }
};
struct FHistogram {
  const u32 *keys;
  u32 *counts;
  GAR_HD void operator()(u32 i) const { GAR_ATOMIC_ADD(&counts[keys[i]], 1u); }
};
struct FGatherTags {
  const u32 *vals, *tags_in;
  u32 *tags_out;
  GAR_HD void operator()(u32 p) const { tags_out[p] = tags_in[vals[p]]; }
};
struct FMarkOrphanValue {
  DevTables T;
  Work W;
  GAR_HD void operator()(u32 v) const { mark_orphan_value(T, W, v); }
};

// --- count / emit.  counts layout: [GA obj: n][GA orphan: nacc][R53 obj: n][orphan alias: nrec][orphan value: nval][total]
struct CountLayout {
  u32 n, nacc, nrec, nval;
  GAR_HD u32 ga_obj(u32 i) const { return i; }
  GAR_HD u32 ga_orph(u32 a) const { return n + a; }
  GAR_HD u32 r53_obj(u32 i) const { return n + nacc + i; }
  GAR_HD u32 base0() const { return 2 * n + nacc; }
  GAR_HD u32 base1() const { return 2 * n + nacc + nrec; }
  GAR_HD u32 total() const { return 2 * n + nacc + nrec + nval; }
};
struct FGaObj {
  DevTables T;
  Work W;
  CountLayout L;
  u32 *counts;     // count pass: written; emit pass: scanned offsets
  gar_op *ops;     // nullptr in the count pass
  u32 *status;
  GAR_HD void operator()(u32 i) const {
    OpSink s{ops ? ops + counts[L.ga_obj(i)] : nullptr, 0};
    u32 st = ga_reconcile(T, W, i, s);
    if (ops) status[i] = st;
    else counts[L.ga_obj(i)] = s.n;
  }
};
struct FGaOrphan {
  DevTables T;
  Work W;
  CountLayout L;
  u32 *counts;
  gar_op *ops;
  GAR_HD void operator()(u32 a) const {
    OpSink s{ops ? ops + counts[L.ga_orph(a)] : nullptr, 0};
    ga_orphan(T, W, a, s);
    if (!ops) counts[L.ga_orph(a)] = s.n;
  }
};
struct FR53Obj {
  DevTables T;
  Work W;
  CountLayout L;
  u32 *counts;
  gar_op *ops;
  u32 *status;
  GAR_HD void operator()(u32 i) const {
    OpSink s{ops ? ops + counts[L.r53_obj(i)] : nullptr, 0};
    u32 st = r53_reconcile(T, W, i, s);
    if (ops) status[i] = st;
    else counts[L.r53_obj(i)] = s.n;
  }
};
// orphan section order: per zone, alias phase (by record row) then metadata phase (by value row)
struct FR53OrphanAlias {
  DevTables T;
  Work W;
  CountLayout L;
  u32 *counts;
  gar_op *ops;
  GAR_HD void operator()(u32 r) const {
    if (!ops) {
      OpSink s{nullptr, 0};
      r53_orphan_alias(T, W, r, s);
      counts[L.base0() + r] = s.n;
      return;
    }
    u32 z = W.rec_zone[r];
    u32 zv = T.a.rec_val_begin[T.a.zone_rec_begin[z]];  // first value row of the zone
    u32 off = counts[L.base0()] + (counts[L.base1() + zv] - counts[L.base1()]) + (counts[L.base0() + r] - counts[L.base0()]);
    OpSink s{ops + off, 0};
    r53_orphan_alias(T, W, r, s);
  }
};
struct FR53OrphanValue {
  DevTables T;
  Work W;
  CountLayout L;
  u32 *counts;
  gar_op *ops;
  GAR_HD void operator()(u32 v) const {
    if (!ops) {
      counts[L.base1() + v] = W.val_orphan[v];
      return;
    }
    if (!W.val_orphan[v]) return;
    u32 r = W.val_rec[v];
    u32 z = W.rec_zone[r];
    u32 rend = T.a.zone_rec_begin[z + 1];  // alias-phase ops of zones 0..z precede
    u32 off = counts[L.base0()] + (counts[L.base0() + rend] - counts[L.base0()]) + (counts[L.base1() + v] - counts[L.base1()]);
    gar_op o;
    o.head = GAR_OP_HEAD(GAR_OP_R53_DELETE_RECORD, GAR_CTRL_R53, 0);
    o.obj = GAR_NONE;
    o.sub = 1;
    o.a0 = z;
    o.a1 = r;
    o.a2 = v;
    ops[off] = o;
  }
};

// ------------------------------------------------------------------ the pipeline

struct DiffCounts {
  u64 n_ops;
  u64 section_begin[GAR_N_SECTIONS + 1];
  u64 n_dports;
  u32 bad_keys;
};

GAR_HD u32 next_pow2(u32 x) {
  u32 p = 1;
  while (p < x) p <<= 1;
  return p;
}
inline int ilog2(u32 p) {
  int b = 0;
  while ((1u << b) < p) b++;
  return b;
}

template <class B>
struct Pipeline {
  B &be;
  DevTables T;
  Work W{};
  explicit Pipeline(B &b, const DevTables &t) : be(b), T(t) {}

  template <class KeyF>
  HashIdx build_index(int slot, u32 nrows, KeyF keyf) {
    u32 nb = next_pow2(nrows < 16 ? 16 : nrows);
    u32 *keys = (u32 *)be.ensure(S_SORT_KEYS, sizeof(u32) * (size_t)(nrows + 1));
    u32 *vals = (u32 *)be.ensure(S_SORT_VALS, sizeof(u32) * (size_t)(nrows + 1));
    u32 *keys2 = (u32 *)be.ensure(S_SORT_KEYS_ALT, sizeof(u32) * (size_t)(nrows + 1));
    u32 *vals2 = (u32 *)be.ensure(S_SORT_VALS_ALT, sizeof(u32) * (size_t)(nrows + 1));
    u32 *tags = (u32 *)be.ensure(S_SORT_TAGS, sizeof(u32) * (size_t)(nrows + 1));
    u32 *begin = (u32 *)be.ensure(slot + 0, sizeof(u32) * (size_t)(nb + 2));
    u32 *row = (u32 *)be.ensure(slot + 1, sizeof(u32) * (size_t)(nrows + 1));
    u32 *tag = (u32 *)be.ensure(slot + 2, sizeof(u32) * (size_t)(nrows + 1));
    keyf.o = IdxOut{keys, vals, tags, nb - 1, nb};
    be.fill32(begin, 0, nb + 2);
    if (nrows) {
      be.for_each("idx_keys", nrows, keyf);
      be.for_each("idx_histogram", nrows, FHistogram{keys, begin});
    }
    be.exclusive_scan(begin, nb + 2);  // begin[b] = #rows with key < b; begin[nb] = #indexed rows
    if (nrows) {
      be.sort_pairs(keys, vals, keys2, vals2, nrows, ilog2(nb) + 1);
      be.for_each("idx_gather", nrows, FGatherTags{vals, tags, tag});
      be.copy32(row, vals, nrows);
    }
    return HashIdx{begin, row, tag, nb - 1};
  }

  // Runs every stage.  `ops_out` is called once the op count is known and must return the device buffer the
  // ops are written to (capacity >= n_ops).
  template <class OpsAlloc>
  int run(DiffCounts *dc, OpsAlloc ops_alloc) {
    const u32 n = T.o.n_objects, nlbi = T.o.n_lbi, nacc = T.a.n_accels, nzone = T.a.n_zones, nrec = T.a.n_records, nval = T.a.n_values;
    W.derived = (u32 *)be.ensure(S_DERIVED, 4 * (size_t)(n + 1));
    u32 *derived_public = (u32 *)be.out_derived(n);
    u8 *oflags = (u8 *)be.ensure(S_OFLAGS, n + 1);
    W.ann_r53 = (gar_str *)be.ensure(S_ANN_R53, 8 * (size_t)(n + 1));
    W.ann_name = (gar_str *)be.ensure(S_ANN_NAME, 8 * (size_t)(n + 1));
    W.ann_tags = (gar_str *)be.ensure(S_ANN_TAGS, 8 * (size_t)(n + 1));
    W.ann_listen = (gar_str *)be.ensure(S_ANN_LISTEN, 8 * (size_t)(n + 1));
    W.dport_begin = (u32 *)be.out_dport_begin(n);
    W.tok_code = (u8 *)be.out_tok_code(nlbi);
    W.tok_name = (gar_str *)be.out_tok_name(nlbi);
    W.tok_region = (gar_str *)be.out_tok_region(nlbi);
    W.acc_flags = (u32 *)be.ensure(S_ACC_FLAGS, 4 * (size_t)(nacc + 1));
    W.acc_owner_key = (gar_str *)be.ensure(S_ACC_OWNER_KEY, 8 * (size_t)(nacc + 1));
    W.acc_owner = (gar_str *)be.ensure(S_ACC_OWNER, 8 * (size_t)(nacc + 1));
    W.acc_thost = (gar_str *)be.ensure(S_ACC_THOST, 8 * (size_t)(nacc + 1));
    W.acc_managed = (gar_str *)be.ensure(S_ACC_MANAGED, 8 * (size_t)(nacc + 1));
    W.rec_zone = (u32 *)be.ensure(S_REC_ZONE, 4 * (size_t)(nrec + 1));
    W.val_rec = (u32 *)be.ensure(S_VAL_REC, 4 * (size_t)(nval + 1));
    W.val_cls = (u8 *)be.ensure(S_VAL_CLS, nval + 1);
    W.val_key = (gar_str *)be.ensure(S_VAL_KEY, 8 * (size_t)(nval + 1));
    W.val_orphan = (u8 *)be.ensure(S_VAL_ORPHAN, nval + 1);
    u32 *errflag = (u32 *)be.ensure(S_ERRFLAG, 16);
    be.fill32(errflag, 0, 4);

    // stage 1: row-local preprocessing
    if (n) be.for_each("classify_objects", n, FClassify{T, W, derived_public, oflags, errflag});
    if (nlbi) be.for_each("tokenise_hostnames", nlbi, FTokenise{T, W});
    if (nacc) be.for_each("digest_accelerators", nacc, FDigestAccel{T, W});
    if (nrec) be.for_each("expand_rec_zone", nrec, FExpand{T.a.zone_rec_begin, nzone, W.rec_zone});
    if (nval) be.for_each("expand_val_rec", nval, FExpand{T.a.rec_val_begin, nrec, W.val_rec});
    if (nval) be.for_each("classify_values", nval, FClassifyValue{T, W});

    // stage 2: listen-ports annotation -> desired port lists (count, scan, write)
    be.fill32(W.dport_begin, 0, (size_t)n + 1);
    if (n) be.for_each("listen_ports_count", n, FJsonCount{T, W});
    be.exclusive_scan(W.dport_begin, n + 1);
    u32 hdr[2] = {0, 0};
    be.download(&hdr[0], W.dport_begin + n, 4);
    be.download(&hdr[1], errflag, 4);
    dc->n_dports = hdr[0];
    dc->bad_keys = hdr[1];
    if (hdr[1]) return GAR_E_INVALID;
    W.dports = (i32 *)be.out_dports(hdr[0]);
    if (n && hdr[0]) be.for_each("listen_ports_write", n, FJsonWrite{T, W});

    // stage 3: hash indexes
    W.ix_lb = build_index(S_IX_LB, T.a.n_lbs, FKeyLb{T, {}});
    W.ix_owner = build_index(S_IX_OWNER, nacc, FKeyOwner{T, W, {}});
    W.ix_thost = build_index(S_IX_THOST, nacc, FKeyThost{T, W, {}});
    W.ix_zone = build_index(S_IX_ZONE, nzone, FKeyZone{T, {}});
    W.ix_val = build_index(S_IX_VAL, nval, FKeyVal{T, W, {}});
    W.ix_alias = build_index(S_IX_ALIAS, nrec, FKeyAlias{T, W, {}});
    W.ix_obj = build_index(S_IX_OBJ, n, FKeyObj{T, W, {}});
    if (nval) be.for_each("mark_orphan_values", nval, FMarkOrphanValue{T, W});
    W.ix_ovn = build_index(S_IX_OVN, nval, FKeyOvn{T, W, {}});

    // stage 4: count ops
    CountLayout L{n, nacc, nrec, nval};
    u32 *counts = (u32 *)be.ensure(S_COUNTS, 4 * (size_t)(L.total() + 2));
    be.fill32(counts, 0, (size_t)L.total() + 1);
    if (n) be.for_each("ga_objects_count", n, FGaObj{T, W, L, counts, nullptr, nullptr});
    if (nacc) be.for_each("ga_orphans_count", nacc, FGaOrphan{T, W, L, counts, nullptr});
    if (n) be.for_each("r53_objects_count", n, FR53Obj{T, W, L, counts, nullptr, nullptr});
    if (nrec) be.for_each("r53_orphan_alias_count", nrec, FR53OrphanAlias{T, W, L, counts, nullptr});
    if (nval) be.for_each("r53_orphan_value_count", nval, FR53OrphanValue{T, W, L, counts, nullptr});
    be.exclusive_scan(counts, L.total() + 1);
    u32 sec[5];
    be.download(&sec[0], counts + 0, 4);
    be.download(&sec[1], counts + L.ga_orph(0), 4);
    be.download(&sec[2], counts + L.r53_obj(0), 4);
    be.download(&sec[3], counts + L.base0(), 4);
    be.download(&sec[4], counts + L.total(), 4);
    for (int k = 0; k < 5; k++) dc->section_begin[k] = sec[k];
    dc->n_ops = sec[4];

    // stage 5: emit
    gar_op *ops = (gar_op *)ops_alloc(dc->n_ops);
    u32 *st_ga = (u32 *)be.out_status_ga(n);
    u32 *st_r53 = (u32 *)be.out_status_r53(n);
    if (n) be.for_each("ga_objects_emit", n, FGaObj{T, W, L, counts, ops, st_ga});
    if (nacc) be.for_each("ga_orphans_emit", nacc, FGaOrphan{T, W, L, counts, ops});
    if (n) be.for_each("r53_objects_emit", n, FR53Obj{T, W, L, counts, ops, st_r53});
    if (nrec) be.for_each("r53_orphan_alias_emit", nrec, FR53OrphanAlias{T, W, L, counts, ops});
    if (nval) be.for_each("r53_orphan_value_emit", nval, FR53OrphanValue{T, W, L, counts, ops});
    return GAR_OK;
  }
};
