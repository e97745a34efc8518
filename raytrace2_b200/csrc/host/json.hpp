// Minimal JSON reader for the scene / settings / camera files (Appendix B of SURVEY.md).
// The reference parses with nlohmann-json (src/Util.cpp:21-32); only the subset of behaviour its loader relies on is
// reproduced here: objects keep every key, numbers remember whether they were written as integers (nlohmann's
// `value("fov", 90)` with an int default truncates a fractional JSON number, src/Serialize.cpp:34), and a parse
// failure yields a null value rather than an exception (Util.cpp:29-31).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace rt2::json {

struct Value;
using Array = std::vector<Value>;
using Object = std::vector<std::pair<std::string, Value>>;

struct Value {
  enum class Kind { kNull, kBool, kNumber, kString, kArray, kObject };
  Kind kind{Kind::kNull};
  bool boolean{false};
  double number{0};
  bool is_integer{false};
  std::string string;
  std::shared_ptr<Array> array;
  std::shared_ptr<Object> object;

  bool IsNull() const { return kind == Kind::kNull; }
  bool IsObject() const { return kind == Kind::kObject; }
  bool IsArray() const { return kind == Kind::kArray; }
  bool IsNumber() const { return kind == Kind::kNumber; }
  bool IsString() const { return kind == Kind::kString; }
  bool IsBool() const { return kind == Kind::kBool; }

  const Value* Find(const char* key) const {
    if (!IsObject()) return nullptr;
    for (const auto& kv : *object)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool Contains(const char* key) const { return Find(key) != nullptr; }
  size_t Size() const { return IsArray() ? array->size() : (IsObject() ? object->size() : 0); }
  const Value& At(size_t i) const { return (*array)[i]; }

  // nlohmann `value(key, default)` equivalents. The arithmetic conversions mirror get<T>() on a JSON number.
  double GetDouble(const char* key, double def) const {
    const Value* v = Find(key);
    return (v && v->IsNumber()) ? v->number : def;
  }
  float GetFloat(const char* key, float def) const {
    const Value* v = Find(key);
    return (v && v->IsNumber()) ? static_cast<float>(v->number) : def;
  }
  // int default => the JSON number is converted to int (truncation toward zero), like nlohmann.
  int GetInt(const char* key, int def) const {
    const Value* v = Find(key);
    return (v && v->IsNumber()) ? static_cast<int>(v->number) : def;
  }
  unsigned GetUint(const char* key, unsigned def) const {
    const Value* v = Find(key);
    return (v && v->IsNumber()) ? static_cast<unsigned>(static_cast<long long>(v->number)) : def;
  }
  bool GetBool(const char* key, bool def) const {
    const Value* v = Find(key);
    if (!v) return def;
    if (v->IsBool()) return v->boolean;
    if (v->IsNumber()) return v->number != 0;
    return def;
  }
  std::string GetString(const char* key, const std::string& def) const {
    const Value* v = Find(key);
    return (v && v->IsString()) ? v->string : def;
  }
  // std::array<real, N> default: every element double -> float.
  template <int N>
  bool GetFloatArray(const char* key, float (&out)[N]) const {
    const Value* v = Find(key);
    if (!v || !v->IsArray() || v->array->size() < static_cast<size_t>(N)) return false;
    for (int i = 0; i < N; i++) {
      if (!(*v->array)[i].IsNumber()) return false;
    }
    for (int i = 0; i < N; i++) out[i] = static_cast<float>((*v->array)[i].number);
    return true;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& text) : s_(text.data()), end_(text.data() + text.size()) {}
  // Returns false on malformed input; `out` is then null.
  bool Parse(Value& out, std::string* err) {
    SkipWs();
    if (!ParseValue(out, 0)) {
      if (err) *err = err_;
      out = Value{};
      return false;
    }
    SkipWs();
    if (s_ != end_) {
      if (err) *err = "trailing characters after JSON value";
      out = Value{};
      return false;
    }
    return true;
  }

 private:
  const char* s_;
  const char* end_;
  std::string err_;

  void SkipWs() {
    while (s_ < end_ && (*s_ == ' ' || *s_ == '\n' || *s_ == '\t' || *s_ == '\r')) s_++;
  }
  bool Fail(const char* msg) {
    err_ = msg;
    return false;
  }
  bool ParseValue(Value& out, int depth) {
    if (depth > 256) return Fail("nesting too deep");
    if (s_ >= end_) return Fail("unexpected end of input");
    switch (*s_) {
      case '{': return ParseObject(out, depth);
      case '[': return ParseArray(out, depth);
      case '"':
        out.kind = Value::Kind::kString;
        return ParseString(out.string);
      case 't':
        if (end_ - s_ >= 4 && std::memcmp(s_, "true", 4) == 0) {
          s_ += 4;
          out.kind = Value::Kind::kBool;
          out.boolean = true;
          return true;
        }
        return Fail("bad literal");
      case 'f':
        if (end_ - s_ >= 5 && std::memcmp(s_, "false", 5) == 0) {
          s_ += 5;
          out.kind = Value::Kind::kBool;
          out.boolean = false;
          return true;
        }
        return Fail("bad literal");
      case 'n':
        if (end_ - s_ >= 4 && std::memcmp(s_, "null", 4) == 0) {
          s_ += 4;
          out.kind = Value::Kind::kNull;
          return true;
        }
        return Fail("bad literal");
      default: return ParseNumber(out);
    }
  }
  bool ParseNumber(Value& out) {
    const char* start = s_;
    bool integer = true;
    if (s_ < end_ && *s_ == '-') s_++;
    if (s_ >= end_ || !(*s_ >= '0' && *s_ <= '9')) return Fail("bad number");
    while (s_ < end_ && *s_ >= '0' && *s_ <= '9') s_++;
    if (s_ < end_ && *s_ == '.') {
      integer = false;
      s_++;
      while (s_ < end_ && *s_ >= '0' && *s_ <= '9') s_++;
    }
    if (s_ < end_ && (*s_ == 'e' || *s_ == 'E')) {
      integer = false;
      s_++;
      if (s_ < end_ && (*s_ == '+' || *s_ == '-')) s_++;
      while (s_ < end_ && *s_ >= '0' && *s_ <= '9') s_++;
    }
    std::string tok(start, s_);
    out.kind = Value::Kind::kNumber;
    out.number = std::strtod(tok.c_str(), nullptr);  // correctly rounded, like nlohmann's strtod path
    out.is_integer = integer;
    return true;
  }
  bool ParseString(std::string& out) {
    s_++;  // opening quote
    out.clear();
    while (s_ < end_ && *s_ != '"') {
      if (*s_ == '\\') {
        s_++;
        if (s_ >= end_) return Fail("bad escape");
        switch (*s_) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'u': {
            if (end_ - s_ < 5) return Fail("bad unicode escape");
            unsigned cp = static_cast<unsigned>(std::strtoul(std::string(s_ + 1, s_ + 5).c_str(), nullptr, 16));
            s_ += 4;
            if (cp < 0x80) {
              out.push_back(static_cast<char>(cp));
            } else if (cp < 0x800) {
              out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
              out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
            } else {
              out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
              out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
              out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
            }
            break;
          }
          default: out.push_back(*s_); break;
        }
        s_++;
      } else {
        out.push_back(*s_++);
      }
    }
    if (s_ >= end_) return Fail("unterminated string");
    s_++;  // closing quote
    return true;
  }
  bool ParseArray(Value& out, int depth) {
    s_++;
    out.kind = Value::Kind::kArray;
    out.array = std::make_shared<Array>();
    SkipWs();
    if (s_ < end_ && *s_ == ']') {
      s_++;
      return true;
    }
    while (true) {
      SkipWs();
      Value v;
      if (!ParseValue(v, depth + 1)) return false;
      out.array->emplace_back(std::move(v));
      SkipWs();
      if (s_ >= end_) return Fail("unterminated array");
      if (*s_ == ',') {
        s_++;
        continue;
      }
      if (*s_ == ']') {
        s_++;
        return true;
      }
      return Fail("expected , or ] in array");
    }
  }
  bool ParseObject(Value& out, int depth) {
    s_++;
    out.kind = Value::Kind::kObject;
    out.object = std::make_shared<Object>();
    SkipWs();
    if (s_ < end_ && *s_ == '}') {
      s_++;
      return true;
    }
    while (true) {
      SkipWs();
      if (s_ >= end_ || *s_ != '"') return Fail("expected string key");
      std::string key;
      if (!ParseString(key)) return false;
      SkipWs();
      if (s_ >= end_ || *s_ != ':') return Fail("expected : after key");
      s_++;
      SkipWs();
      Value v;
      if (!ParseValue(v, depth + 1)) return false;
      // duplicate keys: last one wins, like nlohmann's std::map-backed object
      bool replaced = false;
      for (auto& kv : *out.object) {
        if (kv.first == key) {
          kv.second = std::move(v);
          replaced = true;
          break;
        }
      }
      if (!replaced) out.object->emplace_back(std::move(key), std::move(v));
      SkipWs();
      if (s_ >= end_) return Fail("unterminated object");
      if (*s_ == ',') {
        s_++;
        continue;
      }
      if (*s_ == '}') {
        s_++;
        return true;
      }
      return Fail("expected , or } in object");
    }
  }
};

}  // namespace rt2::json
