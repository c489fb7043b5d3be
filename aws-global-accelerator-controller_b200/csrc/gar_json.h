// gar_json.h — single-pass evaluator for the `alb.ingress.kubernetes.io/listen-ports` annotation.
//
// Replaces json.Unmarshal([]byte(val), &[]IngressPort{}) + the port loop of listenerForIngress
// (reference pkg/cloudprovider/aws/global_accelerator.go:517-542).  The reference maps EVERY json error —
// syntax error (encoding/json checkValid) or type mismatch (UnmarshalTypeError, reported after decoding) —
// to "no ports", so one pass that answers {error | port list} is observably identical to Go's
// validate-then-decode.  What must be exact:
//   * the JSON grammar of encoding/json's scanner (whitespace set, number grammar, string escapes, control
//     characters, literals, trailing data, maxNestingDepth = 10000);
//   * struct-field matching: key is unquoted (escapes, \uXXXX) and compared case-insensitively with Go's
//     simple folding (so U+017F matches 'S'); duplicate keys: last one wins; unknown keys are skipped but validated;
//   * value rules for int64 fields: integer literal in int64 range, or null (no-op); anything else is an error;
//   * array elements: object or null; top level: array or null;
//   * per element HTTP is appended before HTTPS, zero values are skipped, int64 -> int32 truncates.
#pragma once

#include "gar_common.h"

#define GAR_JSON_MAX_DEPTH 10000
#define GAR_JSON_STACK_WORDS ((GAR_JSON_MAX_DEPTH + 31) / 32 + 1)

struct JsonCur {
  const u8 *s;
  u32 n, i;
};

GAR_HD bool js_is_ws(u8 c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }
GAR_HD void js_ws(JsonCur &c) {
  while (c.i < c.n && js_is_ws(c.s[c.i])) c.i++;
}
GAR_HD bool js_is_hex(u8 c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'); }
GAR_HD u32 js_hexv(u8 c) { return c <= '9' ? (u32)(c - '0') : (u32)((c | 0x20) - 'a' + 10); }

// fixed literal (true / false / null) starting at c.i
GAR_HD bool js_literal(JsonCur &c, const char *lit, u32 n) {
  if (c.i + n > c.n) return false;
  for (u32 k = 0; k < n; k++)
    if (c.s[c.i + k] != (u8)lit[k]) return false;
  c.i += n;
  return true;
}

// JSON number grammar.  *is_int is set when there is no fraction and no exponent.
GAR_HD bool js_number(JsonCur &c, bool *is_int) {
  u32 i = c.i, n = c.n;
  const u8 *s = c.s;
  if (i < n && s[i] == '-') i++;
  if (i >= n) return false;
  if (s[i] == '0') {
    i++;
  } else if (s[i] >= '1' && s[i] <= '9') {
    while (i < n && s[i] >= '0' && s[i] <= '9') i++;
  } else {
    return false;
  }
  bool integer = true;
  if (i < n && s[i] == '.') {
    integer = false;
    i++;
    if (i >= n || s[i] < '0' || s[i] > '9') return false;
    while (i < n && s[i] >= '0' && s[i] <= '9') i++;
  }
  if (i < n && (s[i] == 'e' || s[i] == 'E')) {
    integer = false;
    i++;
    if (i < n && (s[i] == '+' || s[i] == '-')) i++;
    if (i >= n || s[i] < '0' || s[i] > '9') return false;
    while (i < n && s[i] >= '0' && s[i] <= '9') i++;
  }
  *is_int = integer;
  c.i = i;
  return true;
}

// String starting at the opening quote.  Validates the syntax; when `field` is non-null also matches the
// unquoted key against the struct tags: *field = 1 for HTTP, 2 for HTTPS, 0 otherwise.
GAR_HD bool js_string(JsonCur &c, int *field) {
  const u8 *s = c.s;
  u32 i = c.i + 1, n = c.n;
  const char want[5] = {'H', 'T', 'T', 'P', 'S'};
  u32 runes = 0;
  bool ok = true;  // every rune so far folds to want[k]
  for (;;) {
    if (i >= n) return false;
    u8 b = s[i];
    if (b == '"') {
      i++;
      break;
    }
    if (b < 0x20) return false;
    u32 r;
    if (b == '\\') {
      if (i + 1 >= n) return false;
      u8 e = s[i + 1];
      if (e == 'u') {
        if (i + 6 > n) return false;
        if (!js_is_hex(s[i + 2]) || !js_is_hex(s[i + 3]) || !js_is_hex(s[i + 4]) || !js_is_hex(s[i + 5])) return false;
        r = js_hexv(s[i + 2]) << 12 | js_hexv(s[i + 3]) << 8 | js_hexv(s[i + 4]) << 4 | js_hexv(s[i + 5]);
        i += 6;
        if (r >= 0xD800 && r < 0xE000) r = 0xFFFD;  // a pair decodes above the BMP, a lone half to U+FFFD: neither can match
      } else if (e == '"' || e == '\\' || e == '/' || e == 'b' || e == 'f' || e == 'n' || e == 'r' || e == 't') {
        r = 0xFFFD;  // these unquote to punctuation / control characters, never to a letter
        i += 2;
      } else {
        return false;
      }
    } else if (b < 0x80) {
      r = b;
      i++;
    } else if (b == 0xC5 && i + 1 < n && s[i + 1] == 0xBF) {
      r = 0x017F;  // LATIN SMALL LETTER LONG S: simple-folds to 'S'
      i += 2;
    } else {
      r = 0xFFFD;  // any other non-ASCII byte sequence: no rune in it folds to H/T/P/S
      i++;
    }
    if (r >= 'a' && r <= 'z') r -= 32;
    else if (r == 0x017F) r = 'S';
    if (runes < 5 && ok && r == (u32)want[runes]) {
    } else {
      ok = false;
    }
    runes++;
  }
  if (field) *field = (ok && runes == 4) ? 1 : (ok && runes == 5) ? 2 : 0;
  c.i = i;
  return true;
}

// Validate and skip one value of any type.  `depth` = nesting depth the value sits in; `stk` holds one bit per
// level (1 = object, 0 = array).
GAR_HD bool js_skip_value(JsonCur &c, u32 depth, u32 *stk) {
  const u32 base = depth;
  for (;;) {
    js_ws(c);
    if (c.i >= c.n) return false;
    u8 ch = c.s[c.i];
    bool need_value = false;
    if (ch == '{') {
      depth++;
      if (depth > GAR_JSON_MAX_DEPTH) return false;
      stk[depth >> 5] |= 1u << (depth & 31);
      c.i++;
      js_ws(c);
      if (c.i < c.n && c.s[c.i] == '}') {
        c.i++;
        depth--;
      } else {
        if (c.i >= c.n || c.s[c.i] != '"' || !js_string(c, nullptr)) return false;
        js_ws(c);
        if (c.i >= c.n || c.s[c.i] != ':') return false;
        c.i++;
        need_value = true;
      }
    } else if (ch == '[') {
      depth++;
      if (depth > GAR_JSON_MAX_DEPTH) return false;
      stk[depth >> 5] &= ~(1u << (depth & 31));
      c.i++;
      js_ws(c);
      if (c.i < c.n && c.s[c.i] == ']') {
        c.i++;
        depth--;
      } else {
        need_value = true;
      }
    } else if (ch == '"') {
      if (!js_string(c, nullptr)) return false;
    } else if (ch == '-' || (ch >= '0' && ch <= '9')) {
      bool isint;
      if (!js_number(c, &isint)) return false;
    } else if (ch == 't') {
      if (!js_literal(c, "true", 4)) return false;
    } else if (ch == 'f') {
      if (!js_literal(c, "false", 5)) return false;
    } else if (ch == 'n') {
      if (!js_literal(c, "null", 4)) return false;
    } else {
      return false;
    }
    if (need_value) continue;
    // a value just ended: close containers / move to the next sibling
    for (;;) {
      if (depth == base) return true;
      js_ws(c);
      if (c.i >= c.n) return false;
      u8 d = c.s[c.i];
      bool in_object = (stk[depth >> 5] >> (depth & 31)) & 1u;
      if (d == ',') {
        c.i++;
        if (in_object) {
          js_ws(c);
          if (c.i >= c.n || c.s[c.i] != '"' || !js_string(c, nullptr)) return false;
          js_ws(c);
          if (c.i >= c.n || c.s[c.i] != ':') return false;
          c.i++;
        }
        break;  // next value
      }
      if (d == (in_object ? '}' : ']')) {
        c.i++;
        depth--;
        continue;
      }
      return false;
    }
  }
}

// Integer literal for an int64 field: already known to start with '-' or a digit.
GAR_HD bool js_int64(JsonCur &c, i64 *out) {
  u32 b = c.i;
  bool isint;
  if (!js_number(c, &isint)) return false;
  if (!isint) return false;  // 80.0 / 1e2: strconv.ParseInt fails -> UnmarshalTypeError
  const u8 *s = c.s;
  u32 i = b;
  bool neg = false;
  if (s[i] == '-') {
    neg = true;
    i++;
  }
  u64 acc = 0;
  for (; i < c.i; i++) {
    u64 d = (u64)(s[i] - '0');
    if (acc > (0xFFFFFFFFFFFFFFFFull - d) / 10) return false;  // beyond uint64: certainly out of int64 range
    acc = acc * 10 + d;
  }
  if (neg) {
    if (acc > 0x8000000000000000ull) return false;
    *out = (i64)(0 - acc);
  } else {
    if (acc > 0x7FFFFFFFFFFFFFFFull) return false;
    *out = (i64)acc;
  }
  return true;
}

// Returns the number of ports, or -1 when Go would report an error (=> the reference uses no ports).
// `out` may be null (count only).  `stk` must hold GAR_JSON_STACK_WORDS words.
GAR_HD int json_listen_ports(Str v, i32 *out, u32 *stk) {
  JsonCur c{v.p, v.n, 0};
  int cnt = 0;
  js_ws(c);
  if (c.i >= c.n) return -1;
  if (c.s[c.i] == 'n') {  // null into a slice: nil, no error
    if (!js_literal(c, "null", 4)) return -1;
    js_ws(c);
    return c.i == c.n ? 0 : -1;
  }
  if (c.s[c.i] != '[') return -1;  // any other top-level value: syntax error or UnmarshalTypeError
  c.i++;
  js_ws(c);
  if (c.i < c.n && c.s[c.i] == ']') {
    c.i++;
  } else {
    for (;;) {
      js_ws(c);
      if (c.i >= c.n) return -1;
      i64 http = 0, https = 0;
      u8 ch = c.s[c.i];
      if (ch == 'n') {
        if (!js_literal(c, "null", 4)) return -1;
      } else if (ch == '{') {
        c.i++;
        js_ws(c);
        if (c.i < c.n && c.s[c.i] == '}') {
          c.i++;
        } else {
          for (;;) {
            js_ws(c);
            int field = 0;
            if (c.i >= c.n || c.s[c.i] != '"' || !js_string(c, &field)) return -1;
            js_ws(c);
            if (c.i >= c.n || c.s[c.i] != ':') return -1;
            c.i++;
            if (field) {
              js_ws(c);
              if (c.i >= c.n) return -1;
              u8 vc = c.s[c.i];
              if (vc == 'n') {
                if (!js_literal(c, "null", 4)) return -1;  // null leaves the field as it is
              } else if (vc == '-' || (vc >= '0' && vc <= '9')) {
                i64 x;
                if (!js_int64(c, &x)) return -1;
                if (field == 1) http = x;
                else https = x;
              } else {
                return -1;  // string / bool / object / array into int64 (or garbage)
              }
            } else {
              if (!js_skip_value(c, 2, stk)) return -1;
            }
            js_ws(c);
            if (c.i >= c.n) return -1;
            if (c.s[c.i] == ',') {
              c.i++;
              continue;
            }
            if (c.s[c.i] == '}') {
              c.i++;
              break;
            }
            return -1;
          }
        }
      } else {
        return -1;  // number / string / bool / array element into a struct
      }
      if (http != 0) {
        if (out) out[cnt] = (i32)http;
        cnt++;
      }
      if (https != 0) {
        if (out) out[cnt] = (i32)https;
        cnt++;
      }
      js_ws(c);
      if (c.i >= c.n) return -1;
      if (c.s[c.i] == ',') {
        c.i++;
        continue;
      }
      if (c.s[c.i] == ']') {
        c.i++;
        break;
      }
      return -1;
    }
  }
  js_ws(c);
  return c.i == c.n ? cnt : -1;
}
