// gar_pipeline.h — the diff as a sequence of data-parallel stages over a Backend.
//
// Backend = how a stage runs: the CUDA backend (gar_engine.cu) launches sm_100a kernels on a stream; the
// host-simulation backend (tests/hostsim) runs the same functors in loops so the CPU-only test tier can
// exercise the device logic.  Stages, functors and their order live here once.
//
// Backend interface:
//   template <class F> void for_each(const char *name, u32 n, const F &f);  // f(i) for i in [0, n)
//   template <class F> void for_each_warp(const char *name, u32 n, const F &f);  // f(i, i < n) for i in [0, roundup(n, 32)):
//                                                                                // full warps, f may use GAR_ANY votes
//   void exclusive_scan(u32 *data, u32 n);                                  // in place
//   void sort_pairs(u32 *keys, u32 *vals, u32 *keys_alt, u32 *vals_alt, u32 n, int bits);  // stable, result in keys/vals
//   void fill32(u32 *p, u32 value, size_t n);
//   void *ensure(int slot, size_t bytes);                                   // scratch buffer `slot`, at least `bytes`
//   void download(void *host_dst, const void *dev_src, size_t bytes);       // blocking
//   void download_start(int id, const void *dev_src, size_t bytes);         // id 0..1: small read-back in flight while later
//   void download_wait(int id, void *host_dst, size_t bytes);               //   stages are queued; wait blocks only if needed
#pragma once

#include <stdio.h>
#include <stdlib.h>

#include "gar_rows.h"

#if defined(__CUDA_ARCH__)
#define GAR_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define GAR_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

// ------------------------------------------------------------------ scratch slots

enum Slot {
  S_DERIVED, S_OFLAGS, S_OKEY_HASH, S_STAGE_GA, S_STAGE_R53, S_ANN_R53, S_ANN_NAME, S_ANN_TAGS, S_ANN_LISTEN, S_DPORT_BEGIN, S_DPORTS,
  S_TOK_CODE, S_TOK_NAME, S_TOK_REGION,
  S_ACC_FLAGS, S_ACC_OWNER_KEY, S_ACC_OWNER, S_ACC_THOST, S_ACC_MANAGED, S_ACC_OWNER_HASH, S_ACC_THOST_HASH, S_REC_NAME_HASH, S_VAL_KEY_HASH,
  S_REC_ZONE, S_VAL_REC, S_VAL_CLS, S_VAL_KEY, S_VAL_ORPHAN, S_VAL_LINK, S_ACC_DIGEST,
  S_R53_MODE, S_R53_ACC, S_R53_ACC_DNS, S_PAIR_BEGIN, S_PAIR_OBJ, S_PAIR_HN, S_PAIR_CODE, S_PAIR_ZONE, S_PAIR_REC,
  S_IX_LB, S_IX_OWNER = S_IX_LB + 3, S_IX_THOST = S_IX_OWNER + 3, S_IX_ZONE = S_IX_THOST + 3, S_IX_VAL = S_IX_ZONE + 3,
  S_IX_ALIAS = S_IX_VAL + 3, S_IX_OBJ = S_IX_ALIAS + 3, S_IX_OVN = S_IX_OBJ + 3,
  S_SORT_KEYS = S_IX_OVN + 3, S_SORT_VALS, S_SORT_KEYS_ALT, S_SORT_VALS_ALT, S_SORT_TAGS,
  S_COUNTS, S_STATUS_GA, S_STATUS_R53, S_OPS, S_ERRFLAG, S_IX_EG, S_IX_EG_ENT, S_IX_EG_PAD, S_ACC_CLAIMED,
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
// record row -> zone row, the hash of the record name (every (zone, name) key is derived from it), and its values: value ->
// record, owner-value classification.  One pass over the record table: a record's name and values are neighbours in the slab,
// so they share DRAM sectors.  A record set with MANY values (one hot TXT name claimed by thousands of owners) is not walked
// by its one thread: it goes on a list that FClassifyBigRecords works through with whole blocks.
constexpr u32 REC_INLINE_VALUES = 32;   // more values than this: a whole block works on the record
constexpr u32 REC_HUGE_VALUES = 4096;   // more than this: the whole grid does
struct FPrepareRecord {
  DevTables T;
  Work W;
  u32 *big, *huge;  // [0] = count, then record rows
  GAR_HD void operator()(u32 r) const {
    u32 lo = 0, hi = T.a.n_zones;
    while (hi - lo > 1) {
      u32 mid = (lo + hi) >> 1;
      if (T.a.zone_rec_begin[mid] <= r) lo = mid;
      else hi = mid;
    }
    W.rec_zone[r] = lo;
    W.rec_name_hash[r] = gar_hash(mkstr(T.a.slab, T.a.rec_name[r]));
    u32 v0 = T.a.rec_val_begin[r], v1 = T.a.rec_val_begin[r + 1];
    if (v1 - v0 > REC_INLINE_VALUES) {
      u32 *list = v1 - v0 > REC_HUGE_VALUES ? huge : big;
#if defined(__CUDA_ARCH__)
      list[1 + atomicAdd(list, 1u)] = r;
#else
      list[1 + list[0]++] = r;
#endif
      return;
    }
    for (u32 v = v0; v < v1; v++) {
      W.val_rec[v] = r;
      classify_value(T, W, v);
    }
  }
};
// fixed grid of BIG_BLOCKS x 256 threads: block b takes big records b, b + BIG_BLOCKS, ... and its threads stride over the
// values; the (few) huge records are strided over by the whole grid
constexpr u32 BIG_BLOCKS = 592;  // 4 x 148 SMs
struct FClassifyBigRecords {
  DevTables T;
  Work W;
  const u32 *big, *huge;
  GAR_HD void operator()(u32 i) const {
    const u32 blk = i >> 8, tid = i & 255u, nbig = big[0], nhuge = huge[0];
    for (u32 k = blk; k < nbig; k += BIG_BLOCKS) {
      u32 r = big[1 + k];
      for (u32 v = T.a.rec_val_begin[r] + tid; v < T.a.rec_val_begin[r + 1]; v += 256) {
        W.val_rec[v] = r;
        classify_value(T, W, v);
      }
    }
    for (u32 k = 0; k < nhuge; k++) {
      u32 r = huge[1 + k];
      for (u32 v = T.a.rec_val_begin[r] + i; v < T.a.rec_val_begin[r + 1]; v += BIG_BLOCKS * 256) {
        W.val_rec[v] = r;
        classify_value(T, W, v);
      }
    }
  }
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

// --- index rows: one functor per index says whether a row is indexed, its key hash and its entry payload
// (payload conventions: gar_rows.h "index probes").  The key pass streams the table once (coalesced) and leaves a
// complete 32-byte entry per row; the build then only moves entries.
struct FRowLb {
  DevTables T;
  GAR_HD bool make(u32 i, u64 *h, IdxEntry *e) const {
    *h = key_hash_lb(mkstr(T.a.slab, T.a.lb_region[i]), mkstr(T.a.slab, T.a.lb_name[i]));
    e->a0 = T.a.lb_state[i];
    e->a1 = 0;
    e->s0 = T.a.lb_name[i];
    e->s1 = T.a.lb_region[i];
    return true;
  }
};
struct FRowOwner {
  DevTables T;
  Work W;
  GAR_HD bool make(u32 i, u64 *h, IdxEntry *e) const {
    u32 fl = W.acc_flags[i];
    *h = W.acc_owner_hash[i];
    e->a0 = fl;
    e->a1 = 0;
    e->s0 = W.acc_owner_key[i];
    e->s1 = 0;
    return (fl & ACC_MINE) && (fl & ACC_OWNER_KEYED) && i < W.acc_guest_from;
  }
};
struct FRowThost {
  DevTables T;
  Work W;
  GAR_HD bool make(u32 i, u64 *h, IdxEntry *e) const {
    *h = W.acc_thost_hash[i];
    e->a0 = e->a1 = 0;
    e->s0 = W.acc_thost[i];
    e->s1 = T.a.acc_dns[i];
    return (W.acc_flags[i] & ACC_MINE) != 0 && (!W.sharded || i >= W.acc_guest_from);
  }
};
struct FRowZone {
  DevTables T;
  GAR_HD bool make(u32 i, u64 *h, IdxEntry *e) const {
    Str zn = mkstr(T.a.slab, T.a.zone_name[i]);
    bool valid = zn.n >= 1 && zn.p[zn.n - 1] == '.';
    *h = valid ? key_hash_str(substr(zn, 0, zn.n - 1)) : 0;
    e->a0 = e->a1 = 0;
    e->s0 = T.a.zone_name[i];
    e->s1 = 0;
    return valid;
  }
};
struct FZoneLenMask {
  DevTables T;
  unsigned long long *mask;  // [4]
  GAR_HD void operator()(u32 z) const {
    Str zn = mkstr(T.a.slab, T.a.zone_name[z]);
    if (zn.n < 1 || zn.p[zn.n - 1] != '.') return;  // not indexed (FRowZone)
    u32 b = zn.n - 1 < 255 ? zn.n - 1 : 255;
#if defined(__CUDA_ARCH__)
    atomicOr(&mask[b >> 6], 1ull << (b & 63));
#else
    mask[b >> 6] |= 1ull << (b & 63);
#endif
  }
};
struct FRowVal {
  DevTables T;
  Work W;
  GAR_HD bool make(u32 v, u64 *h, IdxEntry *e) const {
    u32 cls = W.val_cls[v];
    *h = W.val_key_hash[v];
    if (cls == VAL_NOT_OWNER) return false;
    u32 rec = W.val_rec[v];
    u32 kind = (cls & VAL_OWNER_INGRESS) ? 1u : 0u;
    Str nm = mkstr(T.a.slab, T.a.rec_name[rec]);
    u32 bs = find_byte(nm, 0, '\\') < nm.n ? VALNAME_HAS_BACKSLASH : 0u;
    e->a0 = rec;
    e->a1 = W.rec_zone[rec] | bs | (kind << 31);
    e->s0 = W.val_key[v];
    e->s1 = T.a.rec_name[rec];
    return true;
  }
};
struct FRowAlias {
  DevTables T;
  Work W;
  GAR_HD bool make(u32 r, u64 *h, IdxEntry *e) const {
    *h = key_hash_zoned_h(W.rec_zone[r], W.rec_name_hash[r]);
    e->a0 = W.rec_zone[r];
    e->a1 = T.a.rec_type[r];
    e->s0 = T.a.rec_name[r];
    e->s1 = T.a.rec_alias_dns[r];
    return T.a.rec_has_alias[r] != 0;
  }
};
struct FRowObj {
  DevTables T;
  Work W;
  GAR_HD bool make(u32 i, u64 *h, IdxEntry *e) const {
    gar_str ns = T.o.obj_ns[i];
    *h = W.okey_hash[i];
    e->a0 = T.o.obj_kind[i];
    e->a1 = 0;
    e->s0 = GAR_STR(GAR_STR_OFF(ns), GAR_STR_LEN(ns) + 1 + GAR_STR_LEN(T.o.obj_name[i]));
    e->s1 = 0;
    return !(W.derived[i] & OBJ_KEY_BAD);
  }
};
struct FRowOvn {
  DevTables T;
  Work W;
  GAR_HD bool make(u32 v, u64 *h, IdxEntry *e) const {
    u32 rec = W.val_rec[v];
    *h = key_hash_zoned_h(W.rec_zone[rec], W.rec_name_hash[rec]);
    e->a0 = rec;
    e->a1 = W.rec_zone[rec];
    e->s0 = T.a.rec_name[rec];
    e->s1 = T.a.val_value[v];
    return W.val_orphan[v] != 0;
  }
};

// Index build.  Fast path: (1) idx_rows: per row, entry + bucket key + bucket histogram; (2) scan; (3) idx_place: every
// indexed row drops its entry into its bucket with an atomic cursor (any order); (4) idx_order: buckets with >= 2
// entries are sorted by row id, which makes the result identical to a stable sort.  Buckets larger than
// IDX_SMALL_BUCKET (a hot key / adversarial input) raise `overflow`; the pipeline then rebuilds with the stable
// radix sort (idx_gather).
constexpr u32 IDX_SMALL_BUCKET = 48;
template <class RowF>
struct FIdxRows {
  RowF rowf;
  u32 *keys, *vals;
  IdxEntry *tmp;
  u32 *counts;
  u32 mask, nb;
  GAR_HD void operator()(u32 i) const {
    u64 h = 0;
    IdxEntry e;
    bool valid = rowf.make(i, &h, &e);
    u32 k = valid ? hash_bucket(h, mask) : nb;  // rows that are not indexed carry the sentinel key nb
    keys[i] = k;
    if (vals) vals[i] = i;  // only the radix rebuild sorts (key, row) pairs
    if (valid) {
      e.tag = hash_tag(h);
      e.row = i;
      tmp[i] = e;
      GAR_ATOMIC_ADD(&counts[k], 1u);
    }
  }
};
struct FIdxPlace {
  const u32 *keys;
  u32 *cursor;  // [nb] starts as a copy of begin[]
  const IdxEntry *tmp;
  IdxEntry *ent;
  u32 nb;
  GAR_HD void operator()(u32 i) const {
    u32 k = keys[i];
    if (k == nb) return;
#if defined(__CUDA_ARCH__)
    u32 pos = atomicAdd(&cursor[k], 1u);
#else
    u32 pos = cursor[k]++;
#endif
    ent[pos] = tmp[i];
  }
};
struct FIdxOrder {
  const u32 *begin;
  IdxEntry *ent;
  u32 *overflow;
  GAR_HD void operator()(u32 b) const {
    u32 lo = begin[b], m = begin[b + 1] - lo;
    if (m < 2) return;
    if (m > IDX_SMALL_BUCKET) {
      GAR_ATOMIC_ADD(overflow, 1u);
      return;
    }
    for (u32 k = 1; k < m; k++) {  // insertion sort of 32-byte entries by row id, in place
      IdxEntry x = ent[lo + k];
      u32 j = k;
      while (j > 0 && ent[lo + j - 1].row > x.row) {
        ent[lo + j] = ent[lo + j - 1];
        j--;
      }
      if (j != k) ent[lo + j] = x;
    }
  }
};
struct FIdxGather {  // radix fallback: position p of the stable sorted order holds row vals[p]
  const u32 *vals;
  const u32 *nvalid;
  const IdxEntry *tmp;
  IdxEntry *ent;
  GAR_HD void operator()(u32 p) const {
    if (p < *nvalid) ent[p] = tmp[vals[p]];
  }
};
struct FGather5 {
  const u32 *src;
  u32 idx[5];
  const u32 *extra;
  u32 *dst;
  GAR_HD void operator()(u32 k) const { dst[k] = k < 5 ? src[idx[k]] : *extra; }
};
#define GAR_RETRY_WITH_RADIX 1000  // Pipeline::run: rebuild with force_radix (not an error)
#define GAR_REFUSE_EMPTY_CACHE 1001  // orphan deletes against an empty object table (garecon.h "Orphan sweep precondition")
struct FGatherHeader {
  const u32 *ndports, *errflag;
  u32 *dst;
  GAR_HD void operator()(u32) const {
    dst[0] = *ndports;
    dst[1] = *errflag;
  }
};
// the two per-value joins in one pass (independent probe chains overlap): first alias record under the value's (zone, name),
// and whether the owner the value names is still in the object cache
struct FValueJoins {
  DevTables T;
  Work W;
  GAR_HD void operator()(u32 v) const {
    link_value_alias(T, W, v);
    mark_orphan_value(T, W, v);
  }
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
// Object sections: every object is evaluated ONCE.  The evaluation writes the status word, the op count and
// up to OPS_STAGE_CAP ops into the object's staging slot; after the scan FCompactOps moves staged ops to their
// final position (objects with more ops than the slot holds are re-evaluated straight into the output).
constexpr u32 OPS_STAGE_CAP = 4;
// `rows` (nullable) selects a subset of object rows (incremental mode): slot t evaluates object rows[t]; status, counts
// and staging are indexed by the slot, everything about the object by its row.
struct FGaObj {
  DevTables T;
  Work W;
  u32 *counts;    // [slots] this section's counts
  gar_op *stage;  // [slots * OPS_STAGE_CAP]
  u32 *status;
  const u32 *rows;
  GAR_HD void operator()(u32 t, bool valid) const {
    u32 i = (rows && valid) ? rows[t] : t;
    OpSink s{stage + (size_t)t * OPS_STAGE_CAP, 0, OPS_STAGE_CAP};
    u32 st = ga_reconcile(T, W, i, valid, s);
    if (valid) {
      status[t] = st;
      counts[t] = s.n;
    }
  }
};
struct FR53Obj {
  DevTables T;
  Work W;
  u32 *counts;
  gar_op *stage;
  u32 *status;
  const u32 *rows;
  GAR_HD void operator()(u32 t, bool valid) const {
    u32 i = (rows && valid) ? rows[t] : t;
    OpSink s{stage + (size_t)t * OPS_STAGE_CAP, 0, OPS_STAGE_CAP};
    // objects finished by r53_prepare keep the status it wrote; every other status is produced here
    u32 prev = (valid && W.r53_mode[i] == R53_MODE_DONE) ? status[t] : 0;
    u32 st = r53_combine(T, W, i, t, valid, prev, s);
    if (valid) {
      status[t] = st;
      counts[t] = s.n;
    }
  }
};
struct FR53Prepare {
  DevTables T;
  Work W;
  u32 *status;
  const u32 *rows;
  GAR_HD void operator()(u32 t, bool valid) const { r53_prepare(T, W, (rows && valid) ? rows[t] : t, t, valid, status); }
};
struct FR53FillPairs {
  DevTables T;
  Work W;
  const u32 *rows;
  GAR_HD void operator()(u32 t) const { r53_fill_pairs(T, W, rows ? rows[t] : t, t); }
};
struct FR53Pair {
  DevTables T;
  Work W;
  GAR_HD void operator()(u32 p, bool valid) const { r53_pair(T, W, p, valid); }
};
// incremental mode: processDelete of one work-queue key that left the cache (globalaccelerator/service.go:28-52,
// route53/service.go:29-46): cleanup of everything the key owns, in the reference's order
struct DelKeys {
  const u8 *kind;
  const gar_str *key;  // refs into `slab`
  const u8 *slab;
};
struct FDelKeyGa {
  DevTables T;
  Work W;
  DelKeys D;
  u32 *counts;   // count pass: written; emit pass: scanned
  gar_op *ops;   // nullptr in the count pass
  GAR_HD void operator()(u32 k) const {
    u32 kind = D.kind[k];
    Str key = mkstr(D.slab, D.key[k]);
    OpSink s{ops ? ops + counts[k] : nullptr, 0, 0xFFFFFFFFu};
    OwnerIter it = owner_open(W, key_hash_kinded(kind, key), key);
    for (u32 acc; (acc = owner_next(T, W, kind, it)) != GAR_NONE;) put_delete_chain(T, s, GAR_NONE, 0, acc);
    if (!ops) counts[k] = s.n;
  }
};
struct FDelKeyR53 {
  DevTables T;
  Work W;
  DelKeys D;
  u32 *counts;
  gar_op *ops;
  GAR_HD void operator()(u32 k) const {
    u32 kind = D.kind[k];
    Str key = mkstr(D.slab, D.key[k]);
    OpSink s{ops ? ops + counts[k] : nullptr, 0, 0xFFFFFFFFu};
    Owned ow;
    owned_collect(T, W, key_hash_kinded(kind, key), kind, key, ow);
    r53_cleanup(T, W, GAR_NONE, 0, ow, s);
    if (!ops) counts[k] = s.n;
  }
};
struct FRowKnownEg {
  gar_bindings b;
  GAR_HD bool make(u32 i, u64 *h, IdxEntry *e) const {
    *h = gar_hash(mkstr(b.slab, b.known_eg_arn[i]));
    e->a0 = e->a1 = 0;
    e->s0 = b.known_eg_arn[i];
    e->s1 = 0;
    return true;
  }
};
struct FEgb {
  DevTables T;
  Work W;
  DevBindings B;
  u32 *counts;   // count pass: written; emit pass: scanned
  gar_op *ops;
  u32 *status;
  GAR_HD void operator()(u32 k) const {
    OpSink s{ops ? ops + counts[k] : nullptr, 0, 0xFFFFFFFFu};
    u32 st = egb_reconcile(T, W, B, k, s);
    if (ops) status[k] = st;
    else counts[k] = s.n;
  }
};
struct FGatherDerived {
  const u32 *derived, *rows;
  u32 *out;
  GAR_HD void operator()(u32 t) const { out[t] = derived[rows[t]] & 0xFFu; }
};
struct FCompactOps {
  DevTables T;
  Work W;
  const u32 *scanned;  // exclusive-scanned counts of this section (entry t+1 exists: the layout is contiguous)
  const gar_op *stage;
  gar_op *ops;
  u32 ctrl;
  const u32 *rows;
  GAR_HD void operator()(u32 t, bool valid) const {
    u32 i = (rows && valid) ? rows[t] : t;
    u32 off = 0, c = 0;
    if (valid) {
      off = scanned[t];
      c = scanned[t + 1] - off;
      if (c <= OPS_STAGE_CAP)
        for (u32 k = 0; k < c; k++) ops[off + k] = stage[(size_t)t * OPS_STAGE_CAP + k];
    }
    // objects with more ops than a staging slot holds are re-evaluated straight into the output (warp-uniform:
    // the decide functions vote, so the whole warp enters when any lane needs it)
    bool redo = valid && c > OPS_STAGE_CAP;
    if (GAR_ANY(redo)) {
      OpSink s{ops + off, 0, redo ? 0xFFFFFFFFu : 0u};
      if (ctrl == GAR_CTRL_GA) ga_reconcile(T, W, i, redo, s);
      else r53_reconcile(T, W, i, redo, s);  // the per-object routine: same decisions as prepare/pairs/combine
    }
  }
};
struct FGaOrphan {
  DevTables T;
  Work W;
  CountLayout L;
  u32 *counts;
  gar_op *ops;
  GAR_HD void operator()(u32 a) const {
    if (ops && counts[L.ga_orph(a) + 1] == counts[L.ga_orph(a)]) return;  // emit pass: nothing to write
    OpSink s{ops ? ops + counts[L.ga_orph(a)] : nullptr, 0, 0xFFFFFFFFu};
    ga_orphan(T, W, a, s);
    if (!ops) counts[L.ga_orph(a)] = s.n;
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
      OpSink s{nullptr, 0, 0};
      r53_orphan_alias(T, W, r, s);
      counts[L.base0() + r] = s.n;
      return;
    }
    if (counts[L.base0() + r + 1] == counts[L.base0() + r]) return;
    u32 z = W.rec_zone[r];
    u32 zv = T.a.rec_val_begin[T.a.zone_rec_begin[z]];  // first value row of the zone
    u32 off = counts[L.base0()] + (counts[L.base1() + zv] - counts[L.base1()]) + (counts[L.base0() + r] - counts[L.base0()]);
    OpSink s{ops + off, 0, 0xFFFFFFFFu};
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
  bool force_radix = false;  // set when a bucket exceeded IDX_SMALL_BUCKET in the previous attempt
  explicit Pipeline(B &b, const DevTables &t) : be(b), T(t) {}

  // nb: buckets (power of two).  load: target rows per bucket used to size nb from the row count.
  template <class RowF>
  HashIdx build_index(int slot, u32 nrows, u32 load, RowF rowf, u32 *overflow, bool force_radix) {
    u32 nb = next_pow2(nrows / load < 16 ? 16 : nrows / load);
    u32 *keys = (u32 *)be.ensure(S_SORT_KEYS, sizeof(u32) * (size_t)(nrows + 1));
    u32 *vals = (u32 *)be.ensure(S_SORT_VALS, sizeof(u32) * (size_t)(nrows + 1));
    IdxEntry *tmp = (IdxEntry *)be.ensure(S_SORT_TAGS, sizeof(IdxEntry) * (size_t)(nrows + 1));
    u32 *begin = (u32 *)be.ensure(slot + 0, sizeof(u32) * (size_t)(nb + 2));
    IdxEntry *ent = (IdxEntry *)be.ensure(slot + 1, sizeof(IdxEntry) * (size_t)(nrows + 1));
    be.fill32(begin, 0, nb + 2);
    if (nrows) be.for_each("idx_rows", nrows, FIdxRows<RowF>{rowf, keys, force_radix ? vals : nullptr, tmp, begin, nb - 1, nb});
    be.exclusive_scan(begin, nb + 2);  // begin[b] = #rows with key < b; begin[nb] = #indexed rows
    if (nrows && !force_radix) {
      u32 *cursor = (u32 *)be.ensure(S_SORT_KEYS_ALT, sizeof(u32) * (size_t)(nb + 1));
      be.copy32(cursor, begin, nb);
      be.for_each("idx_place", nrows, FIdxPlace{keys, cursor, tmp, ent, nb});
      be.for_each("idx_order", nb, FIdxOrder{begin, ent, overflow});
    } else if (nrows) {
      u32 *keys2 = (u32 *)be.ensure(S_SORT_KEYS_ALT, sizeof(u32) * (size_t)(nrows + 1));
      u32 *vals2 = (u32 *)be.ensure(S_SORT_VALS_ALT, sizeof(u32) * (size_t)(nrows + 1));
      be.sort_pairs(keys, vals, keys2, vals2, nrows, ilog2(nb) + 1);
      be.for_each("idx_gather", nrows, FIdxGather{vals, begin + nb, tmp, ent});
    }
    return HashIdx{begin, ent, nb - 1};
  }

  // ---- prepare: everything that depends on the snapshot only (stages 1-3).  Runs once per loaded snapshot; the
  // digests and indexes stay resident for every later diff (full or incremental).
  bool prepared = false;
  u64 n_dports = 0;
  u32 *errflag = nullptr;

  u32 acc_guest_from = 0xFFFFFFFFu;  // sharded mode: set before prepare()
  u32 sharded = 0;
  bool orphan_sweep = true;        // GAR_FLAG_NO_ORPHANS clears it
  bool allow_empty_cache = false;  // GAR_FLAG_ALLOW_EMPTY_CACHE

  void alloc_work() {
    const u32 n = T.o.n_objects, nlbi = T.o.n_lbi, nacc = T.a.n_accels, nrec = T.a.n_records, nval = T.a.n_values;
    W.derived = (u32 *)be.ensure(S_DERIVED, 4 * (size_t)(n + 1));
    W.okey_hash = (u64 *)be.ensure(S_OKEY_HASH, 8 * (size_t)(n + 1));
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
    W.acc_owner_hash = (u64 *)be.ensure(S_ACC_OWNER_HASH, 8 * (size_t)(nacc + 1));
    W.acc_thost_hash = (u64 *)be.ensure(S_ACC_THOST_HASH, 8 * (size_t)(nacc + 1));
    W.rec_name_hash = (u64 *)be.ensure(S_REC_NAME_HASH, 8 * (size_t)(nrec + 1));
    W.val_key_hash = (u64 *)be.ensure(S_VAL_KEY_HASH, 8 * (size_t)(nval + 1));
    W.rec_zone = (u32 *)be.ensure(S_REC_ZONE, 4 * (size_t)(nrec + 1));
    W.val_rec = (u32 *)be.ensure(S_VAL_REC, 4 * (size_t)(nval + 1));
    W.val_cls = (u8 *)be.ensure(S_VAL_CLS, nval + 1);
    W.val_key = (gar_str *)be.ensure(S_VAL_KEY, 8 * (size_t)(nval + 1));
    W.val_orphan = (u8 *)be.ensure(S_VAL_ORPHAN, nval + 1);
    W.val_link = (ValLink *)be.ensure(S_VAL_LINK, sizeof(ValLink) * (size_t)(nval + 1));
    W.acc_digest = (AccDigest *)be.ensure(S_ACC_DIGEST, sizeof(AccDigest) * (size_t)(nacc + 1));
    W.r53_mode = (u8 *)be.ensure(S_R53_MODE, (size_t)n + 1);
    W.r53_acc = (u32 *)be.ensure(S_R53_ACC, 4 * (size_t)(n + 1));
    W.r53_acc_dns = (gar_str *)be.ensure(S_R53_ACC_DNS, 8 * (size_t)(n + 1));
    errflag = (u32 *)be.ensure(S_ERRFLAG, 256);
    be.fill32(errflag, 0, 4);
    W.acc_guest_from = acc_guest_from;
    W.sharded = sharded;
    W.acc_claimed = nullptr;
  }
  void stage1() {
    const u32 n = T.o.n_objects, nlbi = T.o.n_lbi, nacc = T.a.n_accels, nrec = T.a.n_records, nval = T.a.n_values;
    u32 *derived_public = (u32 *)be.out_derived(n);
    // stage 1: row-local preprocessing
    if (n) be.for_each("classify_objects", n, FClassify{T, W, derived_public, nullptr, errflag});
    if (nlbi) be.for_each("tokenise_hostnames", nlbi, FTokenise{T, W});
    if (nacc) be.for_each("digest_accelerators", nacc, FDigestAccel{T, W});
    if (nrec) {  // includes the records' values
      u32 *big = (u32 *)be.ensure(S_SORT_KEYS, 4 * (size_t)(nval / REC_INLINE_VALUES + 2));
      u32 *huge = (u32 *)be.ensure(S_SORT_VALS, 4 * (size_t)(nval / REC_HUGE_VALUES + 2));
      be.fill32(big, 0, 1);
      be.fill32(huge, 0, 1);
      be.for_each("prepare_records", nrec, FPrepareRecord{T, W, big, huge});
      if (nval > REC_INLINE_VALUES) be.for_each("prepare_records", BIG_BLOCKS * 256, FClassifyBigRecords{T, W, big, huge});
    }
  }
  // index build that does not assume small buckets: fast per-bucket build first, stable radix rebuild if a bucket overflowed
  template <class RowF>
  HashIdx build_index_checked(int slot, u32 nrows, u32 load, RowF rowf) {
    u32 *overflow = errflag + 1;
    HashIdx ix = build_index(slot, nrows, load, rowf, overflow, force_radix);
    if (!force_radix && nrows) {
      u32 ov = 0;
      be.download(&ov, overflow, 4);
      if (ov) {
        be.fill32(overflow, 0, 1);
        ix = build_index(slot, nrows, load, rowf, overflow, true);
      }
    }
    return ix;
  }
  // what the sharded mode's routing needs of a rank's slice: the row-local pass + the (zone, name) -> alias record index
  void prepare_route() {
    alloc_work();
    stage1();
    W.ix_alias = build_index_checked(S_IX_ALIAS, T.a.n_records, 2, FRowAlias{T, W});
  }
  // a directory shard's tables (gar_shard.h): probes in lbi_*, load balancers, accelerator stubs
  void prepare_directory() {
    alloc_work();
    if (T.o.n_lbi) be.for_each("tokenise_hostnames", T.o.n_lbi, FTokenise{T, W});
    if (T.a.n_accels) be.for_each("digest_accelerators", T.a.n_accels, FDigestAccel{T, W});
    W.ix_lb = build_index_checked(S_IX_LB, T.a.n_lbs, 1, FRowLb{T});
    W.ix_thost = build_index_checked(S_IX_THOST, T.a.n_accels, 1, FRowThost{T, W});
  }

  int prepare() {
    const u32 n = T.o.n_objects, nacc = T.a.n_accels, nzone = T.a.n_zones, nrec = T.a.n_records, nval = T.a.n_values;
    alloc_work();
    stage1();
    // stage 2: listen-ports annotation -> desired port lists (count, scan, write)
    be.fill32(W.dport_begin, 0, (size_t)n + 1);
    if (n) be.for_each("listen_ports_count", n, FJsonCount{T, W});
    be.exclusive_scan(W.dport_begin, n + 1);
    be.for_each("gather_header", 1, FGatherHeader{W.dport_begin + n, errflag, errflag + 4});
    be.download_start(0, errflag + 4, 8);  // port total + layout-rule flag: read back while the indexes are built

    // stage 3: hash indexes
    u32 *overflow = errflag + 1;
    W.ix_lb = build_index(S_IX_LB, T.a.n_lbs, 1, FRowLb{T}, overflow, force_radix);
    W.ix_owner = build_index(S_IX_OWNER, nacc, 1, FRowOwner{T, W}, overflow, force_radix);
    W.ix_thost = build_index(S_IX_THOST, nacc, 1, FRowThost{T, W}, overflow, force_radix);
    W.ix_zone = build_index(S_IX_ZONE, nzone, 1, FRowZone{T}, overflow, force_radix);
    {
      unsigned long long *zm = (unsigned long long *)(errflag + 16 + 8);  // 32 bytes inside the 128-byte flag block, 8-byte aligned
      be.fill32((u32 *)zm, 0, 8);
      if (nzone) be.for_each("zone_len_mask", nzone, FZoneLenMask{T, zm});
      W.zone_len_mask = (const u64 *)zm;
    }
    W.ix_val = build_index(S_IX_VAL, nval, 1, FRowVal{T, W}, overflow, force_radix);
    W.ix_alias = build_index(S_IX_ALIAS, nrec, 2, FRowAlias{T, W}, overflow, force_radix);
    W.ix_obj = build_index(S_IX_OBJ, n, 1, FRowObj{T, W}, overflow, force_radix);
    if (nval) be.for_each("value_joins", nval, FValueJoins{T, W});
    W.ix_ovn = build_index(S_IX_OVN, nval, 8, FRowOvn{T, W}, overflow, force_radix);
    // stage 2, second half: the port lists, now that their total is on the host
    u32 hdr[2] = {0, 0};
    be.download_wait(0, hdr, sizeof(hdr));
    n_dports = hdr[0];
    if (hdr[1]) return GAR_E_INVALID;
    W.dports = (i32 *)be.out_dports(hdr[0]);
    if (n && hdr[0]) be.for_each("listen_ports_write", n, FJsonWrite{T, W});
    return GAR_OK;
  }

  // route53 ensure in relational form over `slots` object slots (rows == nullptr: slot = object row)
  // (begin: per-object filter + pair counts, the pair total starts travelling to the host; end: the pairs themselves.
  //  Stages queued between the two overlap the read-back.)
  void r53_relational_begin(u32 slots, const u32 *rows, u32 *st_r53) {
    W.pair_begin = (u32 *)be.ensure(S_PAIR_BEGIN, 4 * (size_t)(slots + 2));
    be.fill32(W.pair_begin, 0, (size_t)slots + 1);
    if (slots) be.for_each_warp("r53_prepare", slots, FR53Prepare{T, W, st_r53, rows});
    be.exclusive_scan(W.pair_begin, slots + 1);
    be.download_start(1, W.pair_begin + slots, 4);
  }
  u64 n_pairs = 0;  // size of the (object, hostname) relation of the last decide (gar_last_counters)
  void r53_relational_end(u32 slots, const u32 *rows) {
    u32 npairs = 0;
    be.download_wait(1, &npairs, 4);
    n_pairs = npairs;
    W.pair_obj = (u32 *)be.ensure(S_PAIR_OBJ, 4 * (size_t)(npairs + 1));
    W.pair_hn = (gar_str *)be.ensure(S_PAIR_HN, 8 * (size_t)(npairs + 1));
    W.pair_code = (u8 *)be.ensure(S_PAIR_CODE, (size_t)npairs + 1);
    W.pair_zone = (u32 *)be.ensure(S_PAIR_ZONE, 4 * (size_t)(npairs + 1));
    W.pair_rec = (u32 *)be.ensure(S_PAIR_REC, 4 * (size_t)(npairs + 1));
    if (npairs) {
      be.for_each("r53_fill_pairs", slots, FR53FillPairs{T, W, rows});
      be.for_each_warp("r53_pairs", npairs, FR53Pair{T, W});
    }
  }

  // ---- full diff: every object + the orphan sections
  template <class OpsAlloc>
  int decide_all(DiffCounts *dc, OpsAlloc ops_alloc) {
    const u32 n = T.o.n_objects, nacc = T.a.n_accels, nrec = T.a.n_records, nval = T.a.n_values;
    u32 *overflow = errflag + 1;
    // stage 4: evaluate every object once (status + count + staged ops); count the orphan sections
    CountLayout L{n, nacc, nrec, nval};
    u32 *counts = (u32 *)be.ensure(S_COUNTS, 4 * (size_t)(L.total() + 2));
    be.fill32(counts, 0, (size_t)L.total() + 1);
    gar_op *stage_ga = (gar_op *)be.ensure(S_STAGE_GA, sizeof(gar_op) * ((size_t)n * OPS_STAGE_CAP + 1));
    gar_op *stage_r53 = (gar_op *)be.ensure(S_STAGE_R53, sizeof(gar_op) * ((size_t)n * OPS_STAGE_CAP + 1));
    u32 *st_ga = (u32 *)be.out_status_ga(n);
    u32 *st_r53 = (u32 *)be.out_status_r53(n);
    r53_relational_begin(n, nullptr, st_r53);
    W.acc_claimed = (u8 *)be.ensure(S_ACC_CLAIMED, (size_t)nacc + 4);
    be.fill32((u32 *)W.acc_claimed, 0, ((size_t)nacc + 3) / 4);
    if (n) be.for_each_warp("ga_objects", n, FGaObj{T, W, counts + L.ga_obj(0), stage_ga, st_ga, nullptr});
    if (nacc && orphan_sweep) be.for_each("ga_orphans_count", nacc, FGaOrphan{T, W, L, counts, nullptr});
    r53_relational_end(n, nullptr);
    if (n) be.for_each_warp("r53_objects", n, FR53Obj{T, W, counts + L.r53_obj(0), stage_r53, st_r53, nullptr});
    if (nrec && orphan_sweep) be.for_each("r53_orphan_alias_count", nrec, FR53OrphanAlias{T, W, L, counts, nullptr});
    if (nval && orphan_sweep) be.for_each("r53_orphan_value_count", nval, FR53OrphanValue{T, W, L, counts, nullptr});
    be.exclusive_scan(counts, L.total() + 1);
    u32 sec[6];
    u32 *secdev = errflag + 8;  // same small scratch buffer
    be.for_each("gather_section_begins", 6, FGather5{counts, {0, L.ga_orph(0), L.r53_obj(0), L.base0(), L.total()}, overflow, secdev});
    be.download(sec, secdev, sizeof(sec));
    if (sec[5] && !force_radix) return GAR_RETRY_WITH_RADIX;  // an index bucket was too large for the fast build
    // an EMPTY object table with owned resources left: an unsynced informer looks exactly like this, and the orphan sections
    // would delete everything the cluster owns.  (A shard that happens to home no object is fine: the cluster is not empty.)
    if (n == 0 && !sharded && !allow_empty_cache && sec[4] != 0) return GAR_REFUSE_EMPTY_CACHE;
    for (int k = 0; k < 5; k++) dc->section_begin[k] = sec[k];
    dc->n_ops = sec[4];
    dc->n_dports = n_dports;

    // stage 5: move ops to their final, canonical positions
    gar_op *ops = (gar_op *)ops_alloc(dc->n_ops);
    if (n) be.for_each_warp("ga_objects_compact", n, FCompactOps{T, W, counts + L.ga_obj(0), stage_ga, ops, GAR_CTRL_GA, nullptr});
    if (nacc && orphan_sweep) be.for_each("ga_orphans_emit", nacc, FGaOrphan{T, W, L, counts, ops});
    W.acc_claimed = nullptr;  // other decide flavours (incremental, bindings) do not maintain it
    if (n) be.for_each_warp("r53_objects_compact", n, FCompactOps{T, W, counts + L.r53_obj(0), stage_r53, ops, GAR_CTRL_R53, nullptr});
    if (nrec && orphan_sweep) be.for_each("r53_orphan_alias_emit", nrec, FR53OrphanAlias{T, W, L, counts, ops});
    if (nval && orphan_sweep) be.for_each("r53_orphan_value_emit", nval, FR53OrphanValue{T, W, L, counts, ops});
    return GAR_OK;
  }

  // ---- incremental diff: `m` object rows + `nd` deleted keys (device arrays).  counts layout:
  // [GA rows: m][GA deleted keys: nd][R53 rows: m][R53 deleted keys: nd][total]
  template <class OpsAlloc>
  int decide_keys(const u32 *rows, u32 m, DelKeys D, u32 nd, DiffCounts *dc, OpsAlloc ops_alloc) {
    u32 *overflow = errflag + 1;
    const u32 total = 2 * m + 2 * nd;
    u32 *counts = (u32 *)be.ensure(S_COUNTS, 4 * (size_t)(total + 2));
    be.fill32(counts, 0, (size_t)total + 1);
    u32 *c_ga = counts, *c_gad = counts + m, *c_r53 = counts + m + nd, *c_r53d = counts + 2 * m + nd;
    gar_op *stage_ga = (gar_op *)be.ensure(S_STAGE_GA, sizeof(gar_op) * ((size_t)m * OPS_STAGE_CAP + 1));
    gar_op *stage_r53 = (gar_op *)be.ensure(S_STAGE_R53, sizeof(gar_op) * ((size_t)m * OPS_STAGE_CAP + 1));
    u32 *st_ga = (u32 *)be.out_status_ga(m);
    u32 *st_r53 = (u32 *)be.out_status_r53(m);
    u32 *derived_out = (u32 *)be.out_derived_keys(m);
    if (m) {
      be.for_each("gather_derived", m, FGatherDerived{W.derived, rows, derived_out});
      be.for_each_warp("ga_objects", m, FGaObj{T, W, c_ga, stage_ga, st_ga, rows});
    }
    if (nd) be.for_each("ga_deleted_keys_count", nd, FDelKeyGa{T, W, D, c_gad, nullptr});
    r53_relational_begin(m, rows, st_r53);
    r53_relational_end(m, rows);
    if (m) be.for_each_warp("r53_objects", m, FR53Obj{T, W, c_r53, stage_r53, st_r53, rows});
    if (nd) be.for_each("r53_deleted_keys_count", nd, FDelKeyR53{T, W, D, c_r53d, nullptr});
    be.exclusive_scan(counts, total + 1);
    u32 sec[6];
    u32 *secdev = errflag + 8;
    be.for_each("gather_section_begins", 6, FGather5{counts, {0, m, m + nd, 2 * m + nd, total}, overflow, secdev});
    be.download(sec, secdev, sizeof(sec));
    if (sec[5] && !force_radix) return GAR_RETRY_WITH_RADIX;
    for (int k = 0; k < 5; k++) dc->section_begin[k] = sec[k];
    dc->n_ops = sec[4];
    dc->n_dports = 0;
    gar_op *ops = (gar_op *)ops_alloc(dc->n_ops);
    if (m) be.for_each_warp("ga_objects_compact", m, FCompactOps{T, W, c_ga, stage_ga, ops, GAR_CTRL_GA, rows});
    if (nd) be.for_each("ga_deleted_keys_emit", nd, FDelKeyGa{T, W, D, c_gad, ops});
    if (m) be.for_each_warp("r53_objects_compact", m, FCompactOps{T, W, c_r53, stage_r53, ops, GAR_CTRL_R53, rows});
    if (nd) be.for_each("r53_deleted_keys_emit", nd, FDelKeyR53{T, W, D, c_r53d, ops});
    return GAR_OK;
  }

  // ---- EndpointGroupBinding set-diff over `nb` bindings (device pointers in `b`)
  template <class OpsAlloc>
  int decide_bindings(const gar_bindings &b, DiffCounts *dc, OpsAlloc ops_alloc) {
    u32 *overflow = errflag + 1;
    const u32 nb = b.n_bindings;
    DevBindings DB{b, build_index(S_IX_EG, b.n_known_egs, 1, FRowKnownEg{b}, overflow, force_radix)};
    u32 *counts = (u32 *)be.ensure(S_COUNTS, 4 * (size_t)(nb + 2));
    be.fill32(counts, 0, (size_t)nb + 1);
    u32 *st = (u32 *)be.out_status_ga(nb);
    if (nb) be.for_each("egb_count", nb, FEgb{T, W, DB, counts, nullptr, nullptr});
    be.exclusive_scan(counts, nb + 1);
    u32 sec[6];
    u32 *secdev = errflag + 8;
    be.for_each("gather_section_begins", 6, FGather5{counts, {0, nb, nb, nb, nb}, overflow, secdev});
    be.download(sec, secdev, sizeof(sec));
    if (sec[5] && !force_radix) return GAR_RETRY_WITH_RADIX;
    for (int k = 0; k < 5; k++) dc->section_begin[k] = sec[k];
    dc->n_ops = sec[4];
    dc->n_dports = 0;
    gar_op *ops = (gar_op *)ops_alloc(dc->n_ops);
    if (nb) be.for_each("egb_emit", nb, FEgb{T, W, DB, counts, ops, st});
    return GAR_OK;
  }
  template <class OpsAlloc>
  int run_bindings(const gar_bindings &b, DiffCounts *dc, OpsAlloc ops_alloc) {
    return run_with([&] { return decide_bindings(b, dc, ops_alloc); });
  }

  // prepare (once per snapshot; redone with the radix build if a bucket overflowed) + one of the decide flavours
  template <class DecideF>
  int run_with(DecideF decide) {
    for (int attempt = 0; attempt < 2; attempt++) {
      if (!prepared) {
        int rc = prepare();
        if (rc != GAR_OK) return rc;
        prepared = true;
      }
      int rc = decide();
      if (rc != GAR_RETRY_WITH_RADIX) return rc;
      force_radix = true;  // some hash bucket was too large for the per-bucket build: rebuild with the stable radix sort
      prepared = false;
    }
    return GAR_E_STATE;
  }
  template <class OpsAlloc>
  int run(DiffCounts *dc, OpsAlloc ops_alloc) {
    return run_with([&] { return decide_all(dc, ops_alloc); });
  }
  template <class OpsAlloc>
  int run_keys(const u32 *rows, u32 m, DelKeys D, u32 nd, DiffCounts *dc, OpsAlloc ops_alloc) {
    return run_with([&] { return decide_keys(rows, m, D, nd, dc, ops_alloc); });
  }
};
