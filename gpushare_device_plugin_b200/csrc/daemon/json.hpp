// json.hpp — a small JSON DOM reader for the kube API objects the path consumes (PodList, Node, Status).
// RFC 8259 syntax, \uXXXX escapes (incl. surrogate pairs) to UTF-8; numbers keep their source text.
#pragma once

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <initializer_list>
#include <string>
#include <utility>
#include <vector>

namespace json {

struct Value {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  std::string s;  // String value, or the literal text of a Number
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;

  const Value *get(const std::string &key) const {
    if (type != Object) return nullptr;
    for (auto &kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  // obj["a"]["b"]... ; nullptr as soon as something is missing
  const Value *path(std::initializer_list<const char *> keys) const {
    const Value *v = this;
    for (const char *k : keys) {
      if (!v) return nullptr;
      v = v->get(k);
    }
    return v;
  }
  std::string str(const std::string &dflt = "") const { return type == String || type == Number ? s : dflt; }
};

class Parser {
 public:
  explicit Parser(const std::string &text) : p_(text.data()), end_(text.data() + text.size()) {}
  bool parse(Value *out) {
    ws();
    if (!value(out, 0)) return false;
    ws();
    return p_ == end_;
  }

 private:
  void ws() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++;
  }
  bool lit(const char *w) {
    const size_t n = strlen(w);
    if ((size_t)(end_ - p_) < n || memcmp(p_, w, n) != 0) return false;
    p_ += n;
    return true;
  }
  static void utf8(uint32_t cp, std::string *o) {
    if (cp < 0x80) {
      o->push_back((char)cp);
    } else if (cp < 0x800) {
      o->push_back((char)(0xC0 | (cp >> 6)));
      o->push_back((char)(0x80 | (cp & 0x3F)));
    } else if (cp < 0x10000) {
      o->push_back((char)(0xE0 | (cp >> 12)));
      o->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      o->push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      o->push_back((char)(0xF0 | (cp >> 18)));
      o->push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      o->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      o->push_back((char)(0x80 | (cp & 0x3F)));
    }
  }
  bool hex4(uint32_t *v) {
    if (end_ - p_ < 4) return false;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
      const char c = *p_++;
      r <<= 4;
      if (c >= '0' && c <= '9') r |= (uint32_t)(c - '0');
      else if (c >= 'a' && c <= 'f') r |= (uint32_t)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') r |= (uint32_t)(c - 'A' + 10);
      else return false;
    }
    *v = r;
    return true;
  }
  bool string(std::string *o) {
    if (p_ >= end_ || *p_ != '"') return false;
    p_++;
    while (p_ < end_) {
      const char c = *p_++;
      if (c == '"') return true;
      if (c != '\\') {
        o->push_back(c);
        continue;
      }
      if (p_ >= end_) return false;
      const char e = *p_++;
      switch (e) {
        case '"': o->push_back('"'); break;
        case '\\': o->push_back('\\'); break;
        case '/': o->push_back('/'); break;
        case 'b': o->push_back('\b'); break;
        case 'f': o->push_back('\f'); break;
        case 'n': o->push_back('\n'); break;
        case 'r': o->push_back('\r'); break;
        case 't': o->push_back('\t'); break;
        case 'u': {
          uint32_t cp;
          if (!hex4(&cp)) return false;
          if (cp >= 0xD800 && cp <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
            p_ += 2;
            uint32_t lo;
            if (!hex4(&lo)) return false;
            if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8(cp, o);
          break;
        }
        default: return false;
      }
    }
    return false;
  }
  bool value(Value *v, int depth) {
    if (depth > 200 || p_ >= end_) return false;
    const char c = *p_;
    if (c == '{') {
      p_++;
      v->type = Value::Object;
      ws();
      if (p_ < end_ && *p_ == '}') {
        p_++;
        return true;
      }
      for (;;) {
        ws();
        std::string k;
        if (!string(&k)) return false;
        ws();
        if (p_ >= end_ || *p_++ != ':') return false;
        ws();
        v->obj.emplace_back(std::move(k), Value());
        if (!value(&v->obj.back().second, depth + 1)) return false;
        ws();
        if (p_ >= end_) return false;
        if (*p_ == ',') {
          p_++;
          continue;
        }
        if (*p_ == '}') {
          p_++;
          return true;
        }
        return false;
      }
    }
    if (c == '[') {
      p_++;
      v->type = Value::Array;
      ws();
      if (p_ < end_ && *p_ == ']') {
        p_++;
        return true;
      }
      for (;;) {
        ws();
        v->arr.emplace_back();
        if (!value(&v->arr.back(), depth + 1)) return false;
        ws();
        if (p_ >= end_) return false;
        if (*p_ == ',') {
          p_++;
          continue;
        }
        if (*p_ == ']') {
          p_++;
          return true;
        }
        return false;
      }
    }
    if (c == '"') {
      v->type = Value::String;
      return string(&v->s);
    }
    if (lit("true")) {
      v->type = Value::Bool;
      v->b = true;
      return true;
    }
    if (lit("false")) {
      v->type = Value::Bool;
      return true;
    }
    if (lit("null")) return true;
    const char *s = p_;
    if (p_ < end_ && *p_ == '-') p_++;
    while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-'))
      p_++;
    if (p_ == s) return false;
    v->type = Value::Number;
    v->s.assign(s, p_);
    return true;
  }
  const char *p_, *end_;
};

inline bool parse(const std::string &text, Value *out) { return Parser(text).parse(out); }

// Compact serialisation (no whitespace, keys in stored order, numbers as their original text).
inline void dump_string(const std::string &t, std::string *out) {
  out->push_back('"');
  for (unsigned char c : t) {
    if (c == '"' || c == '\\') {
      out->push_back('\\');
      out->push_back((char)c);
    } else if (c < 0x20) {
      char e[8];
      snprintf(e, sizeof e, "\\u%04x", c);
      *out += e;
    } else {
      out->push_back((char)c);
    }
  }
  out->push_back('"');
}

inline void dump(const Value &x, std::string *out) {
  switch (x.type) {
    case Value::Null: *out += "null"; break;
    case Value::Bool: *out += x.b ? "true" : "false"; break;
    case Value::Number: *out += x.s; break;
    case Value::String: dump_string(x.s, out); break;
    case Value::Array:
      out->push_back('[');
      for (size_t i = 0; i < x.arr.size(); i++) {
        if (i) out->push_back(',');
        dump(x.arr[i], out);
      }
      out->push_back(']');
      break;
    case Value::Object:
      out->push_back('{');
      for (size_t i = 0; i < x.obj.size(); i++) {
        if (i) out->push_back(',');
        dump_string(x.obj[i].first, out);
        out->push_back(':');
        dump(x.obj[i].second, out);
      }
      out->push_back('}');
      break;
  }
}

}  // namespace json
