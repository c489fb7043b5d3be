// gar_common.h — types and byte-level primitives shared by every kernel of the diff engine.
//
// All row logic is written as GAR_HD functions so that the very same code can be compiled (a) by nvcc into the
// sm_100a kernels that ship in libgarecon.so and (b) by g++ into tests/hostsim, a debugging build that lets
// the CPU-only CI run the device logic against the oracle.  (b) is a test artefact; the product has no CPU path.
#pragma once

#include <stdint.h>
#include <stddef.h>

#include "../../include/garecon.h"

#if defined(__CUDACC__)
#define GAR_HD __host__ __device__ __forceinline__
#define GAR_D __device__ __forceinline__
#else
#define GAR_HD inline
#define GAR_D inline
#endif

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

#define GAR_SLAB_PAD 32  // bytes of readable zero padding the engine keeps after every slab (wide loads may over-read)

// ------------------------------------------------------------------ string views

struct Str {
  const u8 *p;
  u32 n;
};

GAR_HD Str mkstr(const u8 *slab, gar_str r) { return Str{slab + GAR_STR_OFF(r), GAR_STR_LEN(r)}; }
GAR_HD Str substr(Str s, u32 off, u32 n) { return Str{s.p + off, n}; }

// Unaligned little-endian 8-byte load built from two aligned ones.  May touch up to 15 bytes past p:
// every slab carries GAR_SLAB_PAD bytes of padding and slab bases are 16-byte aligned.
GAR_HD u64 ld64u(const u8 *p) {
  uintptr_t a = (uintptr_t)p;
  const u64 *q = (const u64 *)(a & ~(uintptr_t)7);
  unsigned sh = (unsigned)(a & 7) * 8;
  u64 lo = q[0];
  if (sh == 0) return lo;
  u64 hi = q[1];
  return (lo >> sh) | (hi << (64 - sh));
}
GAR_HD u64 lowmask(u32 nbytes) { return nbytes >= 8 ? ~0ull : ((1ull << (nbytes * 8)) - 1); }

// byte-exact equality, 8 bytes per step
GAR_HD bool streq(Str a, Str b) {
  if (a.n != b.n) return false;
  u32 i = 0;
  for (; i + 8 <= a.n; i += 8)
    if (ld64u(a.p + i) != ld64u(b.p + i)) return false;
  if (i < a.n) {
    u64 m = lowmask(a.n - i);
    if ((ld64u(a.p + i) & m) != (ld64u(b.p + i) & m)) return false;
  }
  return true;
}

// equality with a short literal held in registers/constant space (byte loop; literals are <= 80 bytes)
GAR_HD bool streq_lit(Str a, const char *lit, u32 n) {
  if (a.n != n) return false;
  for (u32 i = 0; i < n; i++)
    if (a.p[i] != (u8)lit[i]) return false;
  return true;
}
#define STREQ_LIT(s, lit) streq_lit((s), (lit), (u32)(sizeof(lit) - 1))

GAR_HD bool has_prefix_lit(Str a, const char *lit, u32 n) {
  if (a.n < n) return false;
  for (u32 i = 0; i < n; i++)
    if (a.p[i] != (u8)lit[i]) return false;
  return true;
}
#define HAS_PREFIX_LIT(s, lit) has_prefix_lit((s), (lit), (u32)(sizeof(lit) - 1))

GAR_HD bool has_suffix_lit(Str a, const char *lit, u32 n) {
  if (a.n < n) return false;
  const u8 *p = a.p + (a.n - n);
  for (u32 i = 0; i < n; i++)
    if (p[i] != (u8)lit[i]) return false;
  return true;
}
#define HAS_SUFFIX_LIT(s, lit) has_suffix_lit((s), (lit), (u32)(sizeof(lit) - 1))

// ------------------------------------------------------------------ xxHash64 (XXH64, seed 0) over contiguous bytes

#define XXP1 0x9E3779B185EBCA87ull
#define XXP2 0xC2B2AE3D27D4EB4Full
#define XXP3 0x165667B19E3779F9ull
#define XXP4 0x85EBCA77C2B2AE63ull
#define XXP5 0x27D4EB2F165667C5ull

GAR_HD u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
GAR_HD u64 xx_round(u64 acc, u64 in) { return rotl64(acc + in * XXP2, 31) * XXP1; }
GAR_HD u64 xx_merge(u64 acc, u64 v) { return (acc ^ xx_round(0, v)) * XXP1 + XXP4; }
GAR_HD u64 xx_avalanche(u64 h) {
  h ^= h >> 33;
  h *= XXP2;
  h ^= h >> 29;
  h *= XXP3;
  h ^= h >> 32;
  return h;
}

GAR_HD u64 xxh64(Str s, u64 seed) {
  const u8 *p = s.p;
  u32 n = s.n, i = 0;
  u64 h;
  if (n >= 32) {
    u64 v1 = seed + XXP1 + XXP2, v2 = seed + XXP2, v3 = seed, v4 = seed - XXP1;
    for (; i + 32 <= n; i += 32) {
      v1 = xx_round(v1, ld64u(p + i));
      v2 = xx_round(v2, ld64u(p + i + 8));
      v3 = xx_round(v3, ld64u(p + i + 16));
      v4 = xx_round(v4, ld64u(p + i + 24));
    }
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xx_merge(h, v1);
    h = xx_merge(h, v2);
    h = xx_merge(h, v3);
    h = xx_merge(h, v4);
  } else {
    h = seed + XXP5;
  }
  h += (u64)n;
  for (; i + 8 <= n; i += 8) {
    h ^= xx_round(0, ld64u(p + i));
    h = rotl64(h, 27) * XXP1 + XXP4;
  }
  if (i + 4 <= n) {
    h ^= (ld64u(p + i) & 0xFFFFFFFFull) * XXP1;
    h = rotl64(h, 23) * XXP2 + XXP3;
    i += 4;
  }
  if (i < n) {
    u64 w = ld64u(p + i);
    for (; i < n; i++) {
      h ^= (w & 0xFF) * XXP5;
      h = rotl64(h, 11) * XXP1;
      w >>= 8;
    }
  }
  return xx_avalanche(h);
}

// combine a small integer (kind, zone row ...) or a second hash into a key hash
GAR_HD u64 hmix(u64 a, u64 b) { return xx_avalanche(a * XXP1 + rotl64(b, 29) * XXP2 + XXP5); }

// ------------------------------------------------------------------ bucketed hash index (read side)
//
// An index is a CSR of buckets: bucket b owns entries [begin[b], begin[b+1]).  Entries of a bucket are
// sorted by build-side row, so walking a bucket yields matching rows in table order — which the change
// set's canonical order needs (duplicates are legal and ordered).  `tag` is the upper half of the 64-bit
// key hash; a full key comparison by the caller always follows a tag hit.

struct HashIdx {
  const u32 *begin;  // [nbuckets + 1]
  const u32 *row;    // [n_entries]
  const u32 *tag;    // [n_entries]
  u32 mask;          // nbuckets - 1 (nbuckets is a power of two)
};

struct Cursor {
  u32 pos, end, tag;
};

GAR_HD u32 hash_bucket(u64 h, u32 mask) { return (u32)h & mask; }
GAR_HD u32 hash_tag(u64 h) { return (u32)(h >> 32); }

GAR_HD Cursor idx_open(const HashIdx &ix, u64 h) {
  u32 b = hash_bucket(h, ix.mask);
  Cursor c;
  c.pos = ix.begin[b];
  c.end = ix.begin[b + 1];
  c.tag = hash_tag(h);
  return c;
}
// next candidate row whose tag matches, or GAR_NONE
GAR_HD u32 idx_next(const HashIdx &ix, Cursor &c) {
  while (c.pos < c.end) {
    u32 p = c.pos++;
    if (ix.tag[p] == c.tag) return ix.row[p];
  }
  return GAR_NONE;
}
