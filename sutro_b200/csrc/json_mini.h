// sutro_b200 — a small JSON reader shared by the host-side C++ (schema compiler, model bundle
// manifest).  Objects keep their members in document order (declaration order of schema
// properties matters); numbers keep their token text (exact decimals for schema bounds).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace sb {
namespace json {

struct JVal {
  enum Type { Null, Bool, Num, Str, Arr, Obj } t = Null;
  bool b = false;
  std::string s;  // Str: decoded UTF-8; Num: the token text
  std::vector<JVal> a;
  std::vector<std::pair<std::string, JVal>> o;  // insertion order (declaration order matters)

  const JVal* get(const std::string& k) const {
    if (t != Obj) return nullptr;
    for (auto& kv : o)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  bool has(const std::string& k) const { return get(k) != nullptr; }
  bool is_null() const { return t == Null; }
};

struct SchemaFail {
  std::string msg;
};
[[noreturn]] inline void fail(const std::string& m) { throw SchemaFail{m}; }

struct JParser {
  const char* p;
  const char* end;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  }
  static void put_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) {
      out += static_cast<char>(cp);
    } else if (cp < 0x800) {
      out += static_cast<char>(0xC0 | (cp >> 6));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    } else if (cp < 0x10000) {
      out += static_cast<char>(0xE0 | (cp >> 12));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    } else {
      out += static_cast<char>(0xF0 | (cp >> 18));
      out += static_cast<char>(0x80 | ((cp >> 12) & 0x3F));
      out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F));
      out += static_cast<char>(0x80 | (cp & 0x3F));
    }
  }
  uint32_t hex4() {
    if (end - p < 4) fail("schema JSON: truncated \\u escape");
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) {
      const char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else fail("schema JSON: bad \\u escape");
    }
    return v;
  }
  std::string str() {
    if (p >= end || *p != '"') fail("schema JSON: expected a string");
    ++p;
    std::string out;
    while (true) {
      if (p >= end) fail("schema JSON: unterminated string");
      const unsigned char c = static_cast<unsigned char>(*p++);
      if (c == '"') break;
      if (c != '\\') {
        out += static_cast<char>(c);
        continue;
      }
      if (p >= end) fail("schema JSON: unterminated escape");
      const char e = *p++;
      switch (e) {
        case '"': out += '"'; break;
        case '\\': out += '\\'; break;
        case '/': out += '/'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'n': out += '\n'; break;
        case 'r': out += '\r'; break;
        case 't': out += '\t'; break;
        case 'u': {
          uint32_t cp = hex4();
          if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
            const char* save = p;
            p += 2;
            const uint32_t lo = hex4();
            if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else p = save;
          }
          put_utf8(out, cp);
          break;
        }
        default: fail("schema JSON: bad escape");
      }
    }
    return out;
  }
  JVal value(int depth = 0) {
    if (depth > 200) fail("schema JSON: nesting too deep");
    ws();
    if (p >= end) fail("schema JSON: unexpected end");
    JVal v;
    const char c = *p;
    if (c == '{') {
      ++p;
      v.t = JVal::Obj;
      ws();
      if (p < end && *p == '}') {
        ++p;
        return v;
      }
      while (true) {
        ws();
        std::string k = str();
        ws();
        if (p >= end || *p != ':') fail("schema JSON: expected ':'");
        ++p;
        JVal x = value(depth + 1);
        bool replaced = false;
        for (auto& kv : v.o)
          if (kv.first == k) {
            kv.second = x;  // duplicate key: last one wins, position of the first (like Python)
            replaced = true;
          }
        if (!replaced) v.o.emplace_back(std::move(k), std::move(x));
        ws();
        if (p < end && *p == ',') {
          ++p;
          continue;
        }
        if (p < end && *p == '}') {
          ++p;
          break;
        }
        fail("schema JSON: expected ',' or '}'");
      }
      return v;
    }
    if (c == '[') {
      ++p;
      v.t = JVal::Arr;
      ws();
      if (p < end && *p == ']') {
        ++p;
        return v;
      }
      while (true) {
        v.a.push_back(value(depth + 1));
        ws();
        if (p < end && *p == ',') {
          ++p;
          continue;
        }
        if (p < end && *p == ']') {
          ++p;
          break;
        }
        fail("schema JSON: expected ',' or ']'");
      }
      return v;
    }
    if (c == '"') {
      v.t = JVal::Str;
      v.s = str();
      return v;
    }
    if (end - p >= 4 && !strncmp(p, "true", 4)) {
      p += 4;
      v.t = JVal::Bool;
      v.b = true;
      return v;
    }
    if (end - p >= 5 && !strncmp(p, "false", 5)) {
      p += 5;
      v.t = JVal::Bool;
      return v;
    }
    if (end - p >= 4 && !strncmp(p, "null", 4)) {
      p += 4;
      return v;
    }
    const char* s0 = p;
    if (p < end && *p == '-') ++p;
    while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' ||
                       *p == '+' || *p == '-'))
      ++p;
    if (p == s0) fail("schema JSON: unexpected character");
    v.t = JVal::Num;
    v.s.assign(s0, p);
    return v;
  }
};

// Python json.dumps(s, ensure_ascii=False)
inline std::string dump_string(const std::string& s) {
  std::string out = "\"";
  for (unsigned char c : s) {
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      case '\b': out += "\\b"; break;
      case '\f': out += "\\f"; break;
      default:
        if (c < 0x20) {
          char buf[8];
          snprintf(buf, sizeof buf, "\\u%04x", c);
          out += buf;
        } else {
          out += static_cast<char>(c);
        }
    }
  }
  return out + "\"";
}

// number token -> the text Python's json.dumps gives for the parsed value
inline std::string dump_number(const std::string& tok) {
  const bool integral = tok.find_first_of(".eE") == std::string::npos;
  if (integral) {
    size_t i = 0;
    const bool neg = !tok.empty() && tok[0] == '-';
    if (neg) i = 1;
    while (i + 1 < tok.size() && tok[i] == '0') ++i;
    std::string digits = tok.substr(i);
    if (digits == "0") return "0";
    return (neg ? "-" : "") + digits;
  }
  const double v = strtod(tok.c_str(), nullptr);
  char buf[64];
  for (int prec = 1; prec <= 17; ++prec) {
    snprintf(buf, sizeof buf, "%.*g", prec, v);
    if (strtod(buf, nullptr) == v) break;
  }
  std::string s = buf;
  if (s.find_first_of("eE") == std::string::npos && s.find('.') == std::string::npos &&
      s.find("inf") == std::string::npos && s.find("nan") == std::string::npos)
    s += ".0";
  return s;
}

// json.dumps(v, separators=(",", ":"), ensure_ascii=False)
inline std::string dump_compact(const JVal& v) {
  switch (v.t) {
    case JVal::Null: return "null";
    case JVal::Bool: return v.b ? "true" : "false";
    case JVal::Num: return dump_number(v.s);
    case JVal::Str: return dump_string(v.s);
    case JVal::Arr: {
      std::string out = "[";
      for (size_t i = 0; i < v.a.size(); ++i) out += (i ? "," : "") + dump_compact(v.a[i]);
      return out + "]";
    }
    default: {
      std::string out = "{";
      for (size_t i = 0; i < v.o.size(); ++i)
        out += (i ? "," : "") + dump_string(v.o[i].first) + ":" + dump_compact(v.o[i].second);
      return out + "}";
    }
  }
}

inline JVal parse(const char* text, size_t len) {
  JParser p{text, text + len};
  JVal v = p.value();
  p.ws();
  if (p.p != p.end) fail("JSON: trailing characters");
  return v;
}

}  // namespace json
}  // namespace sb
