// sutro_b200 — output_schema (JSON Schema text) -> byte DFA, natively.
//
// What the reference hands its service is the schema itself (`payload["json_schema"]`,
// sutro/sdk.py:199, produced by normalize_output_schema, sutro/common.py:152-163); the
// service turns it into constrained decoding.  sb200_schema_compile is that step behind the
// C-ABI, so that a host in any language can pass the JSON text and get the automaton the
// engine's mask/sampler kernels consume (sb200_job.fsm_*).  It restates the core of the Python
// compiler (sutro_b200/schema_fsm.py: JSON Schema -> NFA of byte sets -> subset construction
// -> trim) for the constructs Pydantic emits:
//   objects (declared properties, in declaration order; Dict[str, T] as additionalProperties
//   with propertyNames / min/maxProperties), strings (min/maxLength; formats date, time,
//   date-time, uuid, email, uri, ipv4, duration), integers and numbers (inclusive / exclusive bounds as exact digit
//   automata; multipleOf over bounded ranges), booleans, null, arrays (items, min/maxItems; tuples via prefixItems; uniqueItems
//   over small enumerations), enum / const, anyOf / oneOf, allOf of compatible parts, $ref into
//   $defs (recursion unrolled to a fixed depth), type lists.
// Keywords that would constrain the output and are not handled here (pattern, ...) are an
// ERROR, never ignored — the Python host compiles those.  tests/test_schema_native_cpu.py checks that the
// automata accept exactly the same language as the Python compiler's, schema by schema.
//
// Host-only code (no kernels); compiled by nvcc with the rest of the library.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sutro_b200.h"
#include "json_mini.h"
#include "kernels.h"

namespace sb {
namespace {

using namespace json;

// ------------------------------------------------------------------------------------ decimals
// exact value of a JSON number token: sign * digits * 10^exp10
struct Dec {
  bool neg = false;
  std::string digits = "0";  // no leading zeros (except "0")
  int exp10 = 0;
  bool is_zero() const { return digits == "0"; }
};

Dec parse_dec(const std::string& tok) {
  Dec d;
  size_t i = 0;
  if (i < tok.size() && (tok[i] == '-' || tok[i] == '+')) d.neg = tok[i++] == '-';
  std::string digs;
  int frac = 0;
  bool seen_dot = false;
  for (; i < tok.size(); ++i) {
    const char c = tok[i];
    if (c >= '0' && c <= '9') {
      digs += c;
      if (seen_dot) ++frac;
    } else if (c == '.') {
      seen_dot = true;
    } else {
      break;
    }
  }
  int e = 0;
  if (i < tok.size() && (tok[i] == 'e' || tok[i] == 'E')) e = atoi(tok.c_str() + i + 1);
  size_t nz = 0;
  while (nz + 1 < digs.size() && digs[nz] == '0') ++nz;
  digs = digs.substr(nz);
  if (digs.empty()) digs = "0";
  while (digs.size() > 1 && digs.back() == '0' && frac > 0) {  // normalise trailing zeros
    digs.pop_back();
    --frac;
  }
  d.digits = digs;
  d.exp10 = e - frac;
  if (d.is_zero()) d.neg = false, d.exp10 = 0;
  return d;
}

constexpr long long kBig = 4000000000000000000LL;

// round(d * 10^f) towards +inf (ceil) or -inf (floor); *exact = value was an integer
long long scaled_int(const Dec& d, int f, bool ceil_, bool* exact) {
  const int e = d.exp10 + f;
  *exact = true;
  if (d.is_zero()) return 0;
  long long mag = 0;
  bool rem = false;
  if (e >= 0) {
    if (static_cast<int>(d.digits.size()) + e > 18) return d.neg ? -kBig : kBig;
    for (char c : d.digits) mag = mag * 10 + (c - '0');
    for (int i = 0; i < e; ++i) mag *= 10;
  } else {
    const int keep = static_cast<int>(d.digits.size()) + e;  // digits left of the point
    if (keep > 18) return d.neg ? -kBig : kBig;
    for (int i = 0; i < static_cast<int>(d.digits.size()); ++i) {
      if (i < keep) mag = mag * 10 + (d.digits[i] - '0');
      else if (d.digits[i] != '0') rem = true;
    }
  }
  *exact = !rem;
  long long v = d.neg ? -mag : mag;
  if (rem) {
    if (ceil_ && !d.neg) v += 1;   // positive fraction: ceil goes up
    if (!ceil_ && d.neg) v -= 1;   // negative fraction: floor goes down
  }
  return v;
}

int cmp_dec(const Dec& a, const Dec& b) {  // exact comparison through a common scale
  const int f = std::max(0, std::max(-a.exp10, -b.exp10));
  bool ea, eb;
  const long long x = scaled_int(a, f, false, &ea), y = scaled_int(b, f, false, &eb);
  return x < y ? -1 : (x > y ? 1 : 0);
}

// ------------------------------------------------------------------------------------ NFA
struct Mask {
  uint64_t w[4] = {0, 0, 0, 0};
  void set(int b) { w[b >> 6] |= 1ull << (b & 63); }
  void range(int lo, int hi) {
    for (int b = lo; b <= hi; ++b) set(b);
  }
  bool test(int b) const { return (w[b >> 6] >> (b & 63)) & 1; }
  static Mask of(int b) {
    Mask m;
    m.set(b);
    return m;
  }
  static Mask rng(int lo, int hi) {
    Mask m;
    m.range(lo, hi);
    return m;
  }
  Mask operator|(const Mask& o) const {
    Mask m;
    for (int i = 0; i < 4; ++i) m.w[i] = w[i] | o.w[i];
    return m;
  }
  Mask minus(const Mask& o) const {
    Mask m;
    for (int i = 0; i < 4; ++i) m.w[i] = w[i] & ~o.w[i];
    return m;
  }
};

struct Frag {
  int s = -1, e = -1;
  bool none() const { return s < 0; }
};
const Frag kNone{};

struct Nfa {
  std::vector<std::vector<int>> eps;
  std::vector<std::vector<std::pair<Mask, int>>> tr;
  int neu() {
    eps.emplace_back();
    tr.emplace_back();
    if (eps.size() > 4000000) fail("the schema's automaton is too large");
    return static_cast<int>(eps.size()) - 1;
  }
};

struct Builder {
  Nfa n;
  Frag bset(const Mask& m) {
    const int s = n.neu(), e = n.neu();
    n.tr[s].push_back({m, e});
    return {s, e};
  }
  Frag lit(const std::string& data) {
    int s = n.neu(), cur = s;
    for (unsigned char b : data) {
      const int nx = n.neu();
      n.tr[cur].push_back({Mask::of(b), nx});
      cur = nx;
    }
    return {s, cur};
  }
  Frag seq(std::vector<Frag> fs) {
    std::vector<Frag> v;
    for (auto& f : fs)
      if (!f.none()) v.push_back(f);
    for (size_t i = 0; i + 1 < v.size(); ++i) n.eps[v[i].e].push_back(v[i + 1].s);
    return {v.front().s, v.back().e};
  }
  Frag alt(const std::vector<Frag>& fs) {
    const int s = n.neu(), e = n.neu();
    for (auto& f : fs) {
      n.eps[s].push_back(f.s);
      n.eps[f.e].push_back(e);
    }
    return {s, e};
  }
  Frag alt_or_single(const std::vector<Frag>& fs) { return fs.size() == 1 ? fs[0] : alt(fs); }
  Frag opt(const Frag& f) {
    const int s = n.neu(), e = n.neu();
    n.eps[s].push_back(f.s);
    n.eps[s].push_back(e);
    n.eps[f.e].push_back(e);
    return {s, e};
  }
  Frag rep(const std::function<Frag()>& make, int lo, int hi) {
    const int s = n.neu();
    int cur = s;
    const int e = n.neu();
    if (lo == 0) n.eps[s].push_back(e);
    for (int i = 1; i <= hi; ++i) {
      const Frag f = make();
      n.eps[cur].push_back(f.s);
      cur = f.e;
      if (i >= lo) n.eps[cur].push_back(e);
    }
    return {s, e};
  }
  Frag literals(const std::vector<std::string>& words) {
    const int s = n.neu(), e = n.neu();
    std::map<std::pair<int, unsigned char>, int> trie;
    for (auto& w : words) {
      int cur = s;
      for (unsigned char b : w) {
        auto key = std::make_pair(cur, b);
        auto it = trie.find(key);
        if (it == trie.end()) {
          const int nx = n.neu();
          n.tr[cur].push_back({Mask::of(b), nx});
          it = trie.emplace(key, nx).first;
        }
        cur = it->second;
      }
      n.eps[cur].push_back(e);
    }
    return {s, e};
  }
  // one JSON string character: a code point (UTF-8) or one escape sequence
  Frag json_char() {
    Mask ascii_ok;
    ascii_ok.range(0x20, 0x21);
    ascii_ok.range(0x23, 0x5B);
    ascii_ok.range(0x5D, 0x7F);
    const Mask cont = Mask::rng(0x80, 0xBF);
    const Frag two = seq({bset(Mask::rng(0xC2, 0xDF)), bset(cont)});
    Mask e1;
    e1.range(0xE1, 0xEC);
    e1.range(0xEE, 0xEF);
    const Frag three = alt({seq({bset(Mask::of(0xE0)), bset(Mask::rng(0xA0, 0xBF)), bset(cont)}),
                            seq({bset(e1), bset(cont), bset(cont)}),
                            seq({bset(Mask::of(0xED)), bset(Mask::rng(0x80, 0x9F)), bset(cont)})});
    const Frag four =
        alt({seq({bset(Mask::of(0xF0)), bset(Mask::rng(0x90, 0xBF)), bset(cont), bset(cont)}),
             seq({bset(Mask::rng(0xF1, 0xF3)), bset(cont), bset(cont), bset(cont)}),
             seq({bset(Mask::of(0xF4)), bset(Mask::rng(0x80, 0x8F)), bset(cont), bset(cont)})});
    Mask hexd;
    hexd.range(0x30, 0x39);
    hexd.range(0x41, 0x46);
    hexd.range(0x61, 0x66);
    Mask dd;
    dd.set('d');
    dd.set('D');
    Mask simple;
    for (char c : std::string("\"\\/bfnrt")) simple.set(static_cast<unsigned char>(c));
    // \uXXXX except the surrogate block D800-DFFF
    const Frag esc = seq(
        {bset(Mask::of(0x5C)),
         alt({bset(simple),
              seq({bset(Mask::of('u')),
                   alt({seq({bset(hexd.minus(dd)), bset(hexd), bset(hexd), bset(hexd)}),
                        seq({bset(dd), bset(Mask::rng(0x30, 0x37)), bset(hexd), bset(hexd)})})})})});
    return alt({bset(ascii_ok), two, three, four, esc});
  }
  Frag json_string(int lo, int hi) {
    const Mask q = Mask::of(0x22);
    return seq({bset(q), rep([this] { return json_char(); }, lo, hi), bset(q)});
  }

  // equal-length digit strings d with x <= d <= y; a '.' before the digit at index dot_before
  Frag digits_between(const std::string& x, const std::string& y, int dot_before) {
    const int L = static_cast<int>(x.size());
    const Mask any_d = Mask::rng(0x30, 0x39);
    auto digit = [&](int i, const Mask& m) -> Frag {
      Frag f = bset(m);
      return (dot_before >= 0 && i == dot_before) ? seq({lit("."), f}) : f;
    };
    std::function<Frag(int)> free_, at_least, at_most, between;
    free_ = [&](int i) -> Frag {
      std::vector<Frag> fs;
      for (int k = i; k < L; ++k) fs.push_back(digit(k, any_d));
      return fs.empty() ? kNone : seq(fs);
    };
    at_least = [&](int i) -> Frag {
      if (i == L) return kNone;
      const int d = x[i] - '0';
      std::vector<Frag> alts;
      {
        const Frag a = digit(i, Mask::of(0x30 + d));
        alts.push_back(seq({a, at_least(i + 1)}));
      }
      if (d < 9) {
        const Frag a = digit(i, Mask::rng(0x30 + d + 1, 0x39));
        alts.push_back(seq({a, free_(i + 1)}));
      }
      return alt_or_single(alts);
    };
    at_most = [&](int i) -> Frag {
      if (i == L) return kNone;
      const int d = y[i] - '0';
      std::vector<Frag> alts;
      {
        const Frag a = digit(i, Mask::of(0x30 + d));
        alts.push_back(seq({a, at_most(i + 1)}));
      }
      if (d > 0) {
        const Frag a = digit(i, Mask::rng(0x30, 0x30 + d - 1));
        alts.push_back(seq({a, free_(i + 1)}));
      }
      return alt_or_single(alts);
    };
    between = [&](int i) -> Frag {
      if (i == L) return kNone;
      const int dx = x[i] - '0', dy = y[i] - '0';
      if (dx == dy) {
        const Frag a = digit(i, Mask::of(0x30 + dx));
        return seq({a, between(i + 1)});
      }
      std::vector<Frag> alts;
      {
        const Frag a = digit(i, Mask::of(0x30 + dx));
        alts.push_back(seq({a, at_least(i + 1)}));
      }
      {
        const Frag a = digit(i, Mask::of(0x30 + dy));
        alts.push_back(seq({a, at_most(i + 1)}));
      }
      if (dy - dx > 1) {
        const Frag a = digit(i, Mask::rng(0x30 + dx + 1, 0x30 + dy - 1));
        alts.push_back(seq({a, free_(i + 1)}));
      }
      return alt(alts);
    };
    return between(0);
  }

  static std::string zfill(long long v, int total) {
    std::string s = std::to_string(v);
    if (static_cast<int>(s.size()) < total) s = std::string(total - s.size(), '0') + s;
    return s;
  }
  static long long pow10(int k) {
    long long v = 1;
    for (int i = 0; i < k; ++i) v *= 10;
    return v;
  }
  // decimal texts of k / 10^frac for lo <= k <= hi (0 <= lo)
  Frag scaled_range(long long lo, long long hi, int frac) {
    if (hi < lo) return kNone;
    std::vector<Frag> alts;
    const int t0 = std::max(static_cast<int>(std::to_string(lo).size()), frac + 1);
    const int t1 = std::max(static_cast<int>(std::to_string(hi).size()), frac + 1);
    for (int total = t0; total <= t1; ++total) {
      const long long first = total == frac + 1 ? 0 : pow10(total - 1);
      const long long a = std::max(lo, first), z = std::min(hi, pow10(total) - 1);
      if (a <= z)
        alts.push_back(digits_between(zfill(a, total), zfill(z, total), frac ? total - frac : -1));
    }
    if (alts.empty()) return kNone;
    return alt_or_single(alts);
  }
};

// ------------------------------------------------------------------------------------ DFA
struct Dfa {
  std::vector<int32_t> trans;  // [n, 256]
  std::vector<uint8_t> accept, final_;
  int n = 0;
};

constexpr int kMaxDfaStates = 65535;  // the mask-build kernel's grid.y

Dfa determinise(const Nfa& nfa, int start, int end) {
  const int N = static_cast<int>(nfa.eps.size());
  // byte classes: bytes that no transition mask tells apart
  std::vector<int> cls_of(256);
  std::vector<int> rep_byte;
  {
    std::map<std::vector<char>, int> keys;
    std::vector<const Mask*> masks;
    for (auto& v : nfa.tr)
      for (auto& mt : v) masks.push_back(&mt.first);
    // distinct masks only (the signature of a byte is its membership vector)
    std::vector<Mask> uniq;
    {
      std::set<std::array<uint64_t, 4>> seen;
      for (auto* m : masks) {
        std::array<uint64_t, 4> k{m->w[0], m->w[1], m->w[2], m->w[3]};
        if (seen.insert(k).second) uniq.push_back(*m);
      }
    }
    for (int b = 0; b < 256; ++b) {
      std::vector<char> sig(uniq.size());
      for (size_t i = 0; i < uniq.size(); ++i) sig[i] = uniq[i].test(b);
      auto it = keys.find(sig);
      if (it == keys.end()) {
        it = keys.emplace(sig, static_cast<int>(keys.size())).first;
        rep_byte.push_back(b);
      }
      cls_of[b] = it->second;
    }
  }
  const int n_cls = static_cast<int>(rep_byte.size());
  // per NFA state: class -> targets
  std::vector<std::vector<std::pair<int, int>>> step(N);
  for (int s = 0; s < N; ++s)
    for (auto& mt : nfa.tr[s])
      for (int c = 0; c < n_cls; ++c)
        if (mt.first.test(rep_byte[c])) step[s].push_back({c, mt.second});

  std::vector<char> mark(N, 0);
  auto closure = [&](std::vector<int> st) {
    std::vector<int> stack = st;
    for (int s : st) mark[s] = 1;
    while (!stack.empty()) {
      const int s = stack.back();
      stack.pop_back();
      for (int t : nfa.eps[s])
        if (!mark[t]) {
          mark[t] = 1;
          st.push_back(t);
          stack.push_back(t);
        }
    }
    for (int s : st) mark[s] = 0;
    std::sort(st.begin(), st.end());
    return st;
  };
  std::map<std::vector<int>, int> ids;
  std::vector<std::vector<int>> order;
  std::vector<std::vector<int>> rows;
  order.push_back(closure({start}));
  ids[order[0]] = 0;
  for (size_t i = 0; i < order.size(); ++i) {
    std::vector<std::vector<int>> moved(n_cls);
    for (int s : order[i])
      for (auto& ct : step[s]) moved[ct.first].push_back(ct.second);
    std::vector<int> row(n_cls, -1);
    for (int c = 0; c < n_cls; ++c) {
      if (moved[c].empty()) continue;
      std::sort(moved[c].begin(), moved[c].end());
      moved[c].erase(std::unique(moved[c].begin(), moved[c].end()), moved[c].end());
      std::vector<int> nxt = closure(moved[c]);
      auto it = ids.find(nxt);
      if (it == ids.end()) {
        if (static_cast<int>(order.size()) >= kMaxDfaStates)
          fail("the schema's automaton exceeds 65535 states (counted repeats nested in "
               "alternations blow up when determinised)");
        it = ids.emplace(nxt, static_cast<int>(order.size())).first;
        order.push_back(std::move(nxt));
      }
      row[c] = it->second;
    }
    rows.push_back(std::move(row));
  }
  const int S = static_cast<int>(order.size());
  std::vector<char> acc(S), live(S);
  for (int i = 0; i < S; ++i) live[i] = acc[i] = std::binary_search(order[i].begin(), order[i].end(), end);
  for (bool changed = true; changed;) {  // trim: states that cannot reach acceptance die
    changed = false;
    for (int i = 0; i < S; ++i)
      if (!live[i])
        for (int c = 0; c < n_cls; ++c)
          if (rows[i][c] >= 0 && live[rows[i][c]]) {
            live[i] = 1;
            changed = true;
            break;
          }
  }
  if (!live[0]) fail("schema accepts nothing");
  Dfa d;
  d.n = S;
  d.trans.assign(static_cast<size_t>(S) * 256, -1);
  d.accept.assign(S, 0);
  d.final_.assign(S, 0);
  for (int i = 0; i < S; ++i) {
    bool any = false;
    for (int b = 0; b < 256; ++b) {
      const int t = rows[i][cls_of[b]];
      if (t >= 0 && live[t]) {
        d.trans[static_cast<size_t>(i) * 256 + b] = t;
        any = true;
      }
    }
    d.accept[i] = acc[i];
    d.final_[i] = acc[i] && !any;
  }
  return d;
}

// ------------------------------------------------------------------------------------ schema walk
const std::set<std::string> kAnnotations = {
    "title", "description", "default", "examples", "example", "$schema", "$id", "$comment",
    "$defs", "definitions", "deprecated", "readOnly", "writeOnly", "discriminator",
    "contentEncoding", "contentMediaType", "nullable"};
const std::map<std::string, std::set<std::string>> kKeywords = {
    {"object", {"properties", "required", "additionalProperties", "minProperties",
                "maxProperties", "propertyNames"}},
    {"string", {"minLength", "maxLength", "pattern", "format"}},
    {"integer", {"minimum", "maximum", "exclusiveMinimum", "exclusiveMaximum", "multipleOf"}},
    {"number", {"minimum", "maximum", "exclusiveMinimum", "exclusiveMaximum", "multipleOf"}},
    {"array", {"items", "prefixItems", "minItems", "maxItems", "uniqueItems"}},
    {"boolean", {}},
    {"null", {}}};
// the order Python's dict iteration visits _KEYWORDS (type inference from keywords)
const char* const kTypeOrder[] = {"object", "string", "integer", "number", "array", "boolean", "null"};

struct Compiler {
  const JVal& root;
  sb200_fsm_limits lim;
  Builder b;
  std::vector<std::string> ref_stack;

  Compiler(const JVal& r, const sb200_fsm_limits& l) : root(r), lim(l) {}

  static std::set<std::string> keys_of(const JVal& sch) {
    std::set<std::string> k;
    for (auto& kv : sch.o) k.insert(kv.first);
    return k;
  }
  static std::set<std::string> non_annotation(const JVal& sch) {
    std::set<std::string> k;
    for (auto& kv : sch.o)
      if (!kAnnotations.count(kv.first)) k.insert(kv.first);
    return k;
  }
  static long long as_int(const JVal& v, const char* what) {
    if (v.t != JVal::Num) fail(std::string(what) + " must be a number");
    bool exact;
    return scaled_int(parse_dec(v.s), 0, false, &exact);
  }
  static bool truthy(const JVal* v) {
    if (!v) return false;
    switch (v->t) {
      case JVal::Null: return false;
      case JVal::Bool: return v->b;
      case JVal::Num: return !parse_dec(v->s).is_zero();
      case JVal::Str: return !v->s.empty();
      case JVal::Arr: return !v->a.empty();
      default: return !v->o.empty();
    }
  }

  const JVal& resolve(const std::string& ref) {
    if (ref.rfind("#/", 0) != 0) fail("unsupported $ref '" + ref + "'");
    const JVal* node = &root;
    size_t pos = 2;
    while (pos <= ref.size()) {
      size_t nx = ref.find('/', pos);
      if (nx == std::string::npos) nx = ref.size();
      std::string part = ref.substr(pos, nx - pos);
      for (size_t i; (i = part.find("~1")) != std::string::npos;) part.replace(i, 2, "/");
      for (size_t i; (i = part.find("~0")) != std::string::npos;) part.replace(i, 2, "~");
      const JVal* child = node->get(part);
      if (!child) fail("$ref '" + ref + "' does not resolve");
      node = child;
      pos = nx + 1;
    }
    return *node;
  }

  Frag lits(const std::vector<const JVal*>& values) {
    std::vector<std::string> words;
    for (auto* v : values) words.push_back(dump_compact(*v));
    return b.literals(words);
  }
  Frag lits_words(const std::vector<std::string>& words) { return b.literals(words); }

  // ---- numbers ----
  struct Bounds {
    bool has_lo = false, lo_open = false, has_hi = false, hi_open = false;
    Dec lo, hi;
  };
  static Bounds bounds(const JVal& sch) {
    Bounds r;
    auto dec = [&](const char* k, Dec* out) {
      const JVal* v = sch.get(k);
      if (!v || v->is_null()) return false;
      if (v->t != JVal::Num) fail(std::string(k) + " must be a number");
      *out = parse_dec(v->s);
      return true;
    };
    Dec xlo, xhi;
    r.has_lo = dec("minimum", &r.lo);
    r.has_hi = dec("maximum", &r.hi);
    if (dec("exclusiveMinimum", &xlo) && (!r.has_lo || cmp_dec(xlo, r.lo) >= 0)) {
      r.lo = xlo, r.has_lo = true, r.lo_open = true;
    }
    if (dec("exclusiveMaximum", &xhi) && (!r.has_hi || cmp_dec(xhi, r.hi) <= 0)) {
      r.hi = xhi, r.has_hi = true, r.hi_open = true;
    }
    return r;
  }
  // texts of k / 10^frac for lo <= k <= hi (missing bound = capped by max_int_digits)
  Frag signed_range(bool has_lo, long long lo, bool has_hi, long long hi, int frac) {
    const long long cap = Builder::pow10(lim.max_int_digits + frac) - 1;
    lo = !has_lo ? -cap : std::max(lo, -cap);
    hi = !has_hi ? cap : std::min(hi, cap);
    std::vector<Frag> parts;
    if (hi >= 0) {
      Frag f = b.scaled_range(std::max(lo, 0LL), hi, frac);
      if (!f.none()) parts.push_back(f);
    }
    if (lo < 0) {
      Frag neg = b.scaled_range(std::max(-hi, 1LL), -lo, frac);  // "-0" is not produced
      if (!neg.none()) parts.push_back(b.seq({b.lit("-"), neg}));
    }
    if (parts.empty()) return kNone;
    return b.alt_or_single(parts);
  }
  Frag free_integer_body() {
    const Mask digits = Mask::rng(0x30, 0x39);
    const Frag zero = b.lit("0");
    const Frag lead = b.bset(Mask::rng(0x31, 0x39));
    const Frag more = b.rep([&] { return b.bset(digits); }, 0, lim.max_int_digits - 1);
    return b.alt({zero, b.seq({lead, more})});
  }
  // multipleOf over a bounded range: the multiples are spelled out (schema_fsm.py _multiples)
  Frag multiples(const JVal& sch, const Bounds& bd, bool integral) {
    const JVal* m = sch.get("multipleOf");
    if (m->t != JVal::Num) fail("multipleOf must be a number");
    const Dec step = parse_dec(m->s);
    if (step.neg || step.is_zero()) fail("multipleOf must be positive");
    if (!bd.has_lo || !bd.has_hi) fail("multipleOf needs both a minimum and a maximum here");
    const int f = std::max(0, std::max(-step.exp10, std::max(-bd.lo.exp10, -bd.hi.exp10)));
    if (f > 12) fail("multipleOf: too many fraction digits");
    bool e1, e2, e3;
    const long long L = scaled_int(bd.lo, f, true, &e1), H = scaled_int(bd.hi, f, false, &e2),
                    S = scaled_int(step, f, false, &e3);
    if (L <= -kBig || H >= kBig || S >= kBig || S <= 0) fail("multipleOf: bounds out of range");
    auto floor_div = [](long long a, long long b) {
      long long q = a / b;
      if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
      return q;
    };
    long long k0 = -floor_div(-L, S), k1 = floor_div(H, S);   // ceil(L/S), floor(H/S)
    if (bd.lo_open && k0 * S == L) ++k0;
    if (bd.hi_open && k1 * S == H) --k1;
    if (k1 - k0 + 1 > lim.small_int_range)
      fail("multipleOf over more than " + std::to_string(lim.small_int_range) +
           " values is not supported");
    long long p10 = 1;
    for (int i = 0; i < f; ++i) p10 *= 10;
    std::vector<std::string> words;
    for (long long k = k0; k <= k1; ++k) {
      const long long v = k * S;
      if (integral && v % p10 != 0) continue;
      const long long a = v < 0 ? -v : v;
      std::string text = std::to_string(a / p10);
      if (a % p10 != 0) {
        std::string frac = std::to_string(a % p10);
        frac.insert(0, static_cast<size_t>(f) - frac.size(), '0');
        while (!frac.empty() && frac.back() == '0') frac.pop_back();
        text += "." + frac;
      }
      if (v < 0) text.insert(0, "-");
      words.push_back(text);
    }
    if (words.empty()) fail("numeric range is empty");
    return b.literals(words);
  }

  Frag integer(const JVal& sch) {
    const Bounds bd = bounds(sch);
    if (sch.has("multipleOf") && !sch.get("multipleOf")->is_null()) return multiples(sch, bd, true);
    bool has_lo = bd.has_lo, has_hi = bd.has_hi, exact;
    long long ilo = 0, ihi = 0;
    if (has_lo) {
      ilo = scaled_int(bd.lo, 0, true, &exact);
      if (bd.lo_open && exact) ilo += 1;
    }
    if (has_hi) {
      ihi = scaled_int(bd.hi, 0, false, &exact);
      if (bd.hi_open && exact) ihi -= 1;
    }
    if (has_lo && has_hi) {
      if (ihi < ilo) fail("integer range is empty");
      if (ihi - ilo < lim.small_int_range) {
        std::vector<std::string> words;
        for (long long v = ilo; v <= ihi; ++v) words.push_back(std::to_string(v));
        return b.literals(words);
      }
    }
    if (!has_lo && !has_hi) {
      const Frag sign = b.opt(b.lit("-"));
      return b.seq({sign, free_integer_body()});
    }
    Frag f = signed_range(has_lo, ilo, has_hi, ihi, 0);
    if (f.none()) fail("integer range is empty (within max_int_digits digits)");
    return f;
  }
  Frag number(const JVal& sch) {
    const Bounds bd = bounds(sch);
    if (sch.has("multipleOf") && !sch.get("multipleOf")->is_null()) return multiples(sch, bd, false);
    if (!bd.has_hi && (!bd.has_lo || (bd.lo.is_zero() && !bd.lo_open))) {
      const Mask digits = Mask::rng(0x30, 0x39);
      Frag sign = kNone;
      if (!bd.has_lo) sign = b.opt(b.lit("-"));
      const Frag whole = free_integer_body();
      const Frag frac = b.opt(b.seq({b.lit("."), b.rep([&] { return b.bset(digits); }, 1,
                                                       lim.max_frac_digits)}));
      return b.seq({sign, whole, frac});
    }
    std::vector<Frag> parts;
    for (int f = 0; f <= lim.max_frac_digits; ++f) {
      bool exact;
      long long klo = 0, khi = 0;
      if (bd.has_lo) {
        klo = scaled_int(bd.lo, f, true, &exact);
        if (bd.lo_open && exact) klo += 1;
      }
      if (bd.has_hi) {
        khi = scaled_int(bd.hi, f, false, &exact);
        if (bd.hi_open && exact) khi -= 1;
      }
      if (bd.has_lo && bd.has_hi && khi < klo) continue;
      Frag frag = signed_range(bd.has_lo, klo, bd.has_hi, khi, f);
      if (!frag.none()) parts.push_back(frag);
    }
    if (parts.empty()) fail("numeric range is empty (within max_frac_digits fraction digits)");
    return b.alt_or_single(parts);
  }

  // ---- strings ----
  Frag string(const JVal& sch) {
    const int lo = sch.has("minLength") ? static_cast<int>(as_int(*sch.get("minLength"), "minLength")) : 0;
    const bool has_hi = sch.has("maxLength") && !sch.get("maxLength")->is_null();
    const int hi = has_hi ? static_cast<int>(as_int(*sch.get("maxLength"), "maxLength"))
                          : std::max(lo, lim.max_string_chars);
    if (hi < lo) fail("string length range is empty");
    if (sch.has("pattern") && !sch.get("pattern")->is_null())
      fail("pattern is not supported by the native schema compiler (use the Python host)");
    if (sch.has("format") && !sch.get("format")->is_null()) {
      static const std::set<std::string> plain = {"password", "binary", "byte", "regex", "path",
                                                  "file-path", "directory-path"};
      const JVal* f = sch.get("format");
      if (f->t != JVal::Str) fail("format must be a string");
      if (!plain.count(f->s)) {
        // the formats Pydantic emits for date / time / datetime / UUID fields, as the same
        // sound subsets the Python compiler uses (schema_fsm.py _FORMATS: days stop at 28)
        if (sch.has("minLength") || has_hi)
          fail("format together with minLength / maxLength is not supported by the native schema "
               "compiler (use the Python host)");
        Frag body;
        if (f->s == "date") {
          body = fmt_date();
        } else if (f->s == "time") {
          body = fmt_time();
        } else if (f->s == "date-time") {
          body = b.seq({fmt_date(), b.lit("T"), fmt_time(), b.lit("Z")});
        } else if (f->s == "uuid") {
          body = fmt_uuid();
        } else if (f->s == "email") {   // [a-z0-9]{1,12}@[a-z0-9]{1,12}\.(com|org|net)
          body = b.seq({fmt_label(1), b.lit("@"), fmt_label(1), b.lit("."), fmt_tld()});
        } else if (f->s == "uri") {     // https://[a-z0-9]{1,12}\.(com|org|net)(/[a-z0-9]{0,12})?
          body = b.seq({b.lit("https://"), fmt_label(1), b.lit("."), fmt_tld(),
                        b.opt(b.seq({b.lit("/"), fmt_label(0)}))});
        } else if (f->s == "ipv4") {    // four octets 0..255 without leading zeros
          body = b.seq({fmt_octet(), b.rep([&] { return b.seq({b.lit("."), fmt_octet()}); }, 3, 3)});
        } else if (f->s == "duration") {   // PT(\d{1,2}H)?(\d{1,2}M)?\d{1,2}S
          auto d12 = [&] { return b.rep([&] { return digit(); }, 1, 2); };
          body = b.seq({b.lit("PT"), b.opt(b.seq({d12(), b.lit("H")})),
                        b.opt(b.seq({d12(), b.lit("M")})), d12(), b.lit("S")});
        } else {
          fail("string format '" + f->s + "' is not supported by the native schema compiler (use "
               "the Python host)");
        }
        return b.seq({b.lit("\""), body, b.lit("\"")});
      }
    }
    return b.json_string(lo, hi);
  }
  Frag digit(char lo_c = '0', char hi_c = '9') { return b.bset(Mask::rng(lo_c, hi_c)); }
  Frag fmt_date() {   // [12]\d{3}-(0[1-9]|1[0-2])-(0[1-9]|1\d|2[0-8])
    const Frag month = b.alt({b.seq({b.lit("0"), digit('1', '9')}), b.seq({b.lit("1"), digit('0', '2')})});
    const Frag day = b.alt({b.seq({b.lit("0"), digit('1', '9')}), b.seq({b.lit("1"), digit()}),
                            b.seq({b.lit("2"), digit('0', '8')})});
    return b.seq({digit('1', '2'), digit(), digit(), digit(), b.lit("-"), month, b.lit("-"), day});
  }
  Frag fmt_time() {   // ([01]\d|2[0-3]):[0-5]\d:[0-5]\d
    const Frag hour = b.alt({b.seq({digit('0', '1'), digit()}), b.seq({b.lit("2"), digit('0', '3')})});
    return b.seq({hour, b.lit(":"), digit('0', '5'), digit(), b.lit(":"), digit('0', '5'), digit()});
  }
  Frag fmt_label(int lo_n) {   // [a-z0-9]{lo_n,12}
    const Mask m = Mask::rng('a', 'z') | Mask::rng('0', '9');
    return b.rep([&] { return b.bset(m); }, lo_n, 12);
  }
  Frag fmt_tld() { return b.literals({"com", "org", "net"}); }
  Frag fmt_octet() {   // 25[0-5]|2[0-4]\d|1\d\d|[1-9]?\d
    return b.alt({b.seq({b.lit("25"), digit('0', '5')}), b.seq({b.lit("2"), digit('0', '4'), digit()}),
                  b.seq({b.lit("1"), digit(), digit()}), b.seq({b.opt(digit('1', '9')), digit()})});
  }
  Frag fmt_uuid() {   // 8-4-[1-5]3-[89ab]3-12 lower-case hex
    const Mask hex = Mask::rng('0', '9') | Mask::rng('a', 'f');
    auto hexes = [&](int k) { return b.rep([&] { return b.bset(hex); }, k, k); };
    Mask variant = Mask::of('8') | Mask::of('9') | Mask::of('a') | Mask::of('b');
    return b.seq({hexes(8), b.lit("-"), hexes(4), b.lit("-"), digit('1', '5'), hexes(3), b.lit("-"),
                  b.bset(variant), hexes(3), b.lit("-"), hexes(12)});
  }

  // ---- arrays ----
  Frag array(const JVal& sch) {
    static const JVal kEmptyObj = [] {
      JVal v;
      v.t = JVal::Obj;
      return v;
    }();
    const JVal* items = sch.get("items");
    if (!items) items = &kEmptyObj;
    const int lo = sch.has("minItems") ? static_cast<int>(as_int(*sch.get("minItems"), "minItems")) : 0;
    const bool has_hi = sch.has("maxItems") && !sch.get("maxItems")->is_null();
    if (sch.has("prefixItems") && !sch.get("prefixItems")->is_null())
      return tuple(sch, *sch.get("prefixItems"), *items, lo,
                   has_hi ? static_cast<int>(as_int(*sch.get("maxItems"), "maxItems")) : -1);
    const int hi = has_hi ? static_cast<int>(as_int(*sch.get("maxItems"), "maxItems"))
                          : std::max(lo, lim.max_array_items);
    if (hi < lo) fail("array length range is empty");
    if (truthy(sch.get("uniqueItems"))) return unique(*items, lo, hi);
    if (hi == 0) return b.lit("[]");
    if (items->t == JVal::Bool && !items->b) {
      if (lo > 0) fail("array admits no items but minItems > 0");
      return b.lit("[]");
    }
    Frag first;
    {
      const size_t mark = ref_stack.size();
      try {
        first = node(*items);
      } catch (const SchemaFail&) {
        if (lo > 0) throw;
        ref_stack.resize(mark);
        return b.lit("[]");  // items cannot be expressed (recursion floor): stay empty
      }
    }
    const Frag rest = b.rep([&] { return b.seq({b.lit(","), node(*items)}); }, std::max(lo - 1, 0),
                            hi - 1);
    Frag inner = b.seq({first, rest});
    if (lo == 0) inner = b.opt(inner);
    return b.seq({b.lit("["), inner, b.lit("]")});
  }

  // prefixItems: positional schemas, then `items` for the rest (false = nothing more);
  // hi < 0 = no maxItems (schema_fsm.py _tuple)
  Frag tuple(const JVal& sch, const JVal& prefix, const JVal& items, int lo, int hi) {
    if (prefix.t != JVal::Arr) fail("prefixItems must be an array");
    if (truthy(sch.get("uniqueItems"))) fail("uniqueItems with prefixItems is not supported");
    const bool items_false = items.t == JVal::Bool && !items.b;
    const int np = static_cast<int>(prefix.a.size());
    const int n = hi < 0 ? np : std::min(np, hi);
    if (lo > n && items_false) fail("minItems exceeds the number of prefixItems");
    const int extra_lo = std::max(lo - n, 0);
    const int extra_hi = items_false ? 0 : (hi >= 0 ? hi - n : std::max(extra_lo, 0));
    std::vector<Frag> alts;
    for (int take = std::max(std::min(lo, n), 0); take <= n; ++take) {
      if (take < n && extra_lo > 0) continue;
      std::vector<Frag> parts;
      for (int i = 0; i < take; ++i) {
        if (i) parts.push_back(b.lit(","));
        parts.push_back(node(prefix.a[i]));
      }
      if (take == n && extra_hi > 0) {
        Frag more;
        if (n > 0) {
          more = b.rep([&] { return b.seq({b.lit(","), node(items)}); }, extra_lo, extra_hi);
        } else {
          const Frag first = node(items);
          const Frag rest = b.rep([&] { return b.seq({b.lit(","), node(items)}); },
                                  std::max(extra_lo - 1, 0), extra_hi - 1);
          more = b.seq({first, rest});
          if (extra_lo == 0) more = b.opt(more);
        }
        parts.push_back(more);
      }
      if (parts.empty()) {
        alts.push_back(b.lit("[]"));
      } else {
        std::vector<Frag> all{b.lit("[")};
        all.insert(all.end(), parts.begin(), parts.end());
        all.push_back(b.lit("]"));
        alts.push_back(b.seq(all));
      }
    }
    return b.alt_or_single(alts);
  }

  // uniqueItems needs memory a DFA does not have; small enumerations are spelled out
  // (schema_fsm.py _unique)
  Frag unique(const JVal& items_in, int lo, int hi) {
    const JVal* nd = &items_in;
    if (nd->t == JVal::Obj && nd->has("$ref")) {
      if (nd->get("$ref")->t != JVal::Str) fail("$ref must be a string");
      nd = &resolve(nd->get("$ref")->s);
    }
    std::vector<std::string> values;   // compact JSON text of each admissible item
    bool known = false;
    if (nd->t == JVal::Obj) {
      if (nd->has("enum") && nd->get("enum")->t == JVal::Arr) {
        for (auto& v : nd->get("enum")->a) values.push_back(dump_compact(v));
        known = true;
      } else if (nd->has("const")) {
        values.push_back(dump_compact(*nd->get("const")));
        known = true;
      } else if (nd->has("type") && nd->get("type")->t == JVal::Str && nd->get("type")->s == "boolean") {
        values = {"true", "false"};
        known = true;
      }
    }
    if (!known || values.size() > 6)
      fail("uniqueItems is only supported for arrays over at most 6 enumerated values");
    std::vector<std::string> distinct;
    for (auto& v : values)
      if (std::find(distinct.begin(), distinct.end(), v) == distinct.end()) distinct.push_back(v);
    std::vector<std::string> arrays;
    const int top = std::min<int>(hi, static_cast<int>(distinct.size()));
    std::vector<int> pick;
    std::vector<bool> used(distinct.size(), false);
    std::function<void(int)> rec = [&](int k) {
      if (static_cast<int>(pick.size()) == k) {
        std::string a = "[";
        for (size_t i = 0; i < pick.size(); ++i) a += (i ? "," : "") + distinct[pick[i]];
        arrays.push_back(a + "]");
        return;
      }
      for (size_t i = 0; i < distinct.size(); ++i) {
        if (used[i]) continue;
        used[i] = true;
        pick.push_back(static_cast<int>(i));
        rec(k);
        pick.pop_back();
        used[i] = false;
      }
    };
    for (int k = lo; k <= top; ++k) rec(k);
    if (arrays.empty()) fail("uniqueItems: no array satisfies the length bounds");
    return b.literals(arrays);
  }

  // Dict[str, T]: free keys (propertyNames applies), values of one schema; duplicate keys are
  // not excluded (schema_fsm.py _mapping)
  Frag mapping(const JVal& sch, const JVal& value_schema) {
    const int lo = sch.has("minProperties") ? static_cast<int>(as_int(*sch.get("minProperties"), "minProperties")) : 0;
    const bool has_hi = sch.has("maxProperties") && !sch.get("maxProperties")->is_null();
    const int hi = has_hi ? static_cast<int>(as_int(*sch.get("maxProperties"), "maxProperties"))
                          : std::max(lo, lim.max_array_items);
    if (hi < lo) fail("property count range is empty");
    JVal names;
    names.t = JVal::Obj;
    if (const JVal* pn = sch.get("propertyNames"))
      if (pn->t == JVal::Obj) names = *pn;
    auto set_key = [&](const char* k, JVal v) {
      for (auto& kv : names.o)
        if (kv.first == k) return;
      names.o.emplace_back(k, std::move(v));
    };
    JVal str_t;
    str_t.t = JVal::Str;
    str_t.s = "string";
    set_key("type", str_t);
    if (names.get("type")->t != JVal::Str || names.get("type")->s != "string")
      fail("propertyNames must describe strings");
    if (!names.has("minLength") && !names.has("pattern") && !names.has("enum") &&
        !names.has("format") && !names.has("const")) {
      JVal one;      // keep keys non-empty unless the schema says otherwise
      one.t = JVal::Num;
      one.s = "1";
      set_key("minLength", one);
    }
    if (hi == 0) return b.lit("{}");
    auto entry = [&] { return b.seq({node(names), b.lit(":"), node(value_schema)}); };
    const Frag first = entry();
    const Frag rest = b.rep([&] { return b.seq({b.lit(","), entry()}); }, std::max(lo - 1, 0), hi - 1);
    Frag inner = b.seq({first, rest});
    if (lo == 0) inner = b.opt(inner);
    return b.seq({b.lit("{"), inner, b.lit("}")});
  }

  // ---- objects ----
  Frag obj(const JVal& sch) {
    const JVal* props = sch.get("properties");
    const JVal* addl = sch.get("additionalProperties");
    const bool no_props = !props || props->t != JVal::Obj || props->o.empty();
    const long long min_p = sch.has("minProperties") ? as_int(*sch.get("minProperties"), "minProperties") : 0;
    if (no_props) {
      const bool addl_dict = addl && addl->t == JVal::Obj;
      const bool addl_true = addl && addl->t == JVal::Bool && addl->b;
      if (addl_dict || (addl_true && (min_p != 0 || truthy(sch.get("propertyNames"))))) {
        static const JVal kAny = [] {
          JVal v;
          v.t = JVal::Obj;
          return v;
        }();
        return mapping(sch, addl_dict ? *addl : kAny);
      }
      if (min_p > 0) fail("minProperties > 0 on an object without properties");
      return b.lit("{}");
    }
    const long long np = static_cast<long long>(props->o.size());
    if (min_p > np || (sch.has("maxProperties") && !sch.get("maxProperties")->is_null() &&
                       as_int(*sch.get("maxProperties"), "maxProperties") < np))
      fail("min/maxProperties conflict with the listed properties");
    std::vector<Frag> parts;
    parts.push_back(b.lit("{"));
    for (size_t i = 0; i < props->o.size(); ++i) {
      parts.push_back(b.lit(std::string(i ? "," : "") + dump_string(props->o[i].first) + ":"));
      parts.push_back(node(props->o[i].second));
    }
    parts.push_back(b.lit("}"));
    return b.seq(parts);
  }

  Frag any_value() {
    static const JVal kEmptyObj = [] {
      JVal v;
      v.t = JVal::Obj;
      return v;
    }();
    const Frag s = string(kEmptyObj);
    const Frag nmb = number(kEmptyObj);
    const Frag l = lits_words({"true", "false", "null"});
    return b.alt({s, nmb, l});
  }

  Frag node(const JVal& sch) {
    if (sch.t == JVal::Bool && sch.b) return any_value();
    if (sch.t != JVal::Obj) fail("unsupported schema node " + dump_compact(sch));
    if (non_annotation(sch).empty()) return any_value();
    if (const JVal* ref = sch.get("$ref")) {
      if (ref->t != JVal::Str) fail("$ref must be a string");
      auto rest = non_annotation(sch);
      rest.erase("$ref");
      if (!rest.empty()) fail("keywords next to $ref are not supported");
      if (std::count(ref_stack.begin(), ref_stack.end(), ref->s) > lim.max_recursion)
        fail("recursion through " + ref->s + " is deeper than max_recursion levels and nothing "
             "encloses it that could stop (an array with minItems 0, a union)");
      ref_stack.push_back(ref->s);
      struct Pop {
        std::vector<std::string>& st;
        size_t n;
        ~Pop() { st.resize(n); }
      } pop{ref_stack, ref_stack.size() - 1};
      return node(resolve(ref->s));
    }
    if (const JVal* c = sch.get("const")) return lits({c});
    if (const JVal* e = sch.get("enum")) {
      if (e->t != JVal::Arr || e->a.empty()) fail("enum is empty");
      std::vector<const JVal*> vs;
      for (auto& v : e->a) vs.push_back(&v);
      return lits(vs);
    }
    for (const char* k : {"anyOf", "oneOf"}) {
      const JVal* u = sch.get(k);
      if (!u) continue;
      auto rest = non_annotation(sch);
      rest.erase(k);
      if (!rest.empty()) fail(std::string("keywords next to ") + k + " are not supported");
      if (u->t != JVal::Arr) fail(std::string(k) + " must be an array");
      // an alternative this compiler cannot express is left out: the automaton then accepts a
      // subset of the union, which keeps every output valid
      std::vector<Frag> alts;
      std::string errors;
      for (auto& x : u->a) {
        const size_t mark = ref_stack.size();
        try {
          alts.push_back(node(x));
        } catch (const SchemaFail& e) {
          errors += (errors.empty() ? "" : "; ") + e.msg;
          ref_stack.resize(mark);
        }
      }
      if (alts.empty()) fail(std::string("no alternative of ") + k + " is supported: " + errors);
      return b.alt_or_single(alts);
    }
    if (const JVal* all = sch.get("allOf")) {
      if (all->t != JVal::Arr) fail("allOf must be an array");
      JVal rest;
      rest.t = JVal::Obj;
      for (auto& kv : sch.o)
        if (kv.first != "allOf") rest.o.push_back(kv);
      if (all->a.size() == 1 && non_annotation(rest).empty()) return node(all->a[0]);
      JVal merged = rest;
      for (auto& part0 : all->a) {
        const JVal* part = &part0;
        if (part->t == JVal::Obj && part->has("$ref") && all->a.size() > 1) {
          if (part->get("$ref")->t != JVal::Str) fail("$ref must be a string");
          part = &resolve(part->get("$ref")->s);
        }
        if (part->t != JVal::Obj) fail("allOf parts must be schema objects");
        for (auto& kv : part->o) {
          if (kAnnotations.count(kv.first)) continue;
          const JVal* have = merged.get(kv.first);
          if (have && dump_compact(*have) != dump_compact(kv.second))
            fail("allOf parts disagree on '" + kv.first + "': intersections of different "
                 "constraints are not supported");
          if (!have) merged.o.push_back(kv);
        }
      }
      return node(merged);
    }
    for (const char* bad : {"not", "if", "then", "else", "contains", "patternProperties",
                            "dependentRequired", "dependentSchemas", "unevaluatedProperties",
                            "unevaluatedItems"})
      if (sch.has(bad)) fail(std::string("unsupported keyword '") + bad + "'");
    const JVal* tv = sch.get("type");
    std::string t;
    if (tv && tv->t == JVal::Arr) {
      // every member type takes the keywords that apply to it
      std::set<std::string> typed;
      for (auto& kv : kKeywords) typed.insert(kv.second.begin(), kv.second.end());
      std::vector<Frag> alts;
      for (auto& x : tv->a) {
        if (x.t != JVal::Str || !kKeywords.count(x.s)) fail("unsupported type " + dump_compact(x));
        JVal keep;
        keep.t = JVal::Obj;
        for (auto& kv : sch.o) {
          if (kv.first == "type") continue;
          if (!typed.count(kv.first) || kKeywords.at(x.s).count(kv.first)) keep.o.push_back(kv);
        }
        JVal ty;
        ty.t = JVal::Str;
        ty.s = x.s;
        keep.o.emplace_back("type", ty);
        alts.push_back(node(keep));
      }
      if (alts.empty()) fail("type list is empty");
      return b.alt_or_single(alts);
    }
    if (tv && tv->t == JVal::Str) t = tv->s;
    if (!tv || tv->is_null()) {
      const auto keys = keys_of(sch);
      for (const char* name : kTypeOrder) {
        bool hit = false;
        for (auto& k : kKeywords.at(name))
          if (keys.count(k)) hit = true;
        if (hit) {
          t = std::string(name) == "integer" ? "number" : name;
          break;
        }
      }
      if (t.empty()) fail("unsupported schema node " + dump_compact(sch));
    }
    if (!kKeywords.count(t)) fail("unsupported type '" + t + "'");
    {
      std::string extra;
      for (auto& kv : sch.o)
        if (!kAnnotations.count(kv.first) && !kKeywords.at(t).count(kv.first) && kv.first != "type")
          extra += (extra.empty() ? "" : ", ") + kv.first;
      if (!extra.empty())
        fail("unsupported keyword(s) [" + extra + "] on a schema of type '" + t +
             "' (they would constrain the output and cannot be ignored)");
    }
    if (t == "object") return obj(sch);
    if (t == "string") return string(sch);
    if (t == "integer") return integer(sch);
    if (t == "number") return number(sch);
    if (t == "boolean") return lits_words({"true", "false"});
    if (t == "null") return lits_words({"null"});
    return array(sch);
  }
};

struct Schema {
  Dfa dfa;
};

// longest accepted string in bytes, -1 when the language is unbounded (a cycle)
int64_t longest_path(const Dfa& d) {
  const int n = d.n;
  std::vector<std::vector<int>> succ(n);
  for (int s = 0; s < n; ++s) {
    std::set<int> u;
    for (int b = 0; b < 256; ++b)
      if (d.trans[static_cast<size_t>(s) * 256 + b] >= 0) u.insert(d.trans[static_cast<size_t>(s) * 256 + b]);
    succ[s].assign(u.begin(), u.end());
  }
  std::vector<int64_t> depth(n, -1);
  std::vector<char> state(n, 0);
  std::vector<std::pair<int, size_t>> stack{{0, 0}};
  while (!stack.empty()) {
    auto [s, i] = stack.back();
    stack.pop_back();
    if (i == 0) {
      if (state[s] == 2) continue;
      state[s] = 1;
    }
    if (i < succ[s].size()) {
      stack.push_back({s, i + 1});
      const int t = succ[s][i];
      if (state[t] == 1) return -1;
      if (state[t] == 0) stack.push_back({t, 0});
    } else {
      int64_t best = 0;
      for (int t : succ[s]) best = std::max(best, 1 + depth[t]);
      depth[s] = best;
      state[s] = 2;
    }
  }
  return depth[0];
}

}  // namespace
}  // namespace sb

using namespace sb;

extern "C" {

void sb200_fsm_limits_default(sb200_fsm_limits* l) {
  if (!l) return;
  l->max_string_chars = 64;
  l->max_array_items = 8;
  l->max_int_digits = 9;
  l->max_frac_digits = 4;
  l->small_int_range = 2048;
  l->max_recursion = 2;
}

int sb200_schema_compile(const char* json_utf8, int64_t len, const sb200_fsm_limits* limits,
                         void** out) {
  if (out) *out = nullptr;
  if (!json_utf8 || len < 0 || !out) {
    set_last_error("schema_compile: null argument");
    return -1;
  }
  sb200_fsm_limits lim;
  sb200_fsm_limits_default(&lim);
  if (limits) lim = *limits;
  try {
    const JVal root = parse(json_utf8, static_cast<size_t>(len));
    if (root.t != JVal::Obj) fail("schema must be a JSON object");
    Compiler comp(root, lim);
    const Frag f = comp.node(root);
    auto* s = new Schema();
    s->dfa = determinise(comp.b.n, f.s, f.e);
    *out = s;
    return 0;
  } catch (const SchemaFail& e) {
    set_last_error("output_schema: %s", e.msg.c_str());
    return -2;  // argument error (the SDK's ValueError convention)
  } catch (const std::exception& e) {
    set_last_error("schema_compile: %s", e.what());
    return -1;
  }
}

void sb200_schema_destroy(void* schema) { delete static_cast<Schema*>(schema); }

int sb200_schema_tables(void* schema, const int32_t** trans, const uint8_t** accept,
                        const uint8_t** final_states, int* n_states, int* start) {
  if (!schema) {
    set_last_error("schema_tables: null schema");
    return -1;
  }
  const Dfa& d = static_cast<Schema*>(schema)->dfa;
  if (trans) *trans = d.trans.data();
  if (accept) *accept = d.accept.data();
  if (final_states) *final_states = d.final_.data();
  if (n_states) *n_states = d.n;
  if (start) *start = 0;
  return 0;
}

int64_t sb200_schema_longest_path(void* schema) {
  if (!schema) return -1;
  return longest_path(static_cast<Schema*>(schema)->dfa);
}

}  // extern "C"
