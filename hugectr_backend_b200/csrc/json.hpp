// Minimal JSON DOM used for ps.json, the Triton backend-config message and config.pbtxt-as-JSON.
// (The reference uses TritonJson/rapidjson, which are not in this image.)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace hpsx {
namespace json {

class Value {
 public:
  enum Type { Null, Bool, Number, String, Array, Object };

  Value() = default;

  Type type() const { return type_; }
  bool is_null() const { return type_ == Null; }
  bool is_bool() const { return type_ == Bool; }
  bool is_number() const { return type_ == Number; }
  bool is_string() const { return type_ == String; }
  bool is_array() const { return type_ == Array; }
  bool is_object() const { return type_ == Object; }

  bool as_bool() const { return bool_; }
  double as_double() const { return num_; }
  bool is_integer() const { return is_number() && integral_; }
  int64_t as_int() const { return integral_ ? int_ : static_cast<int64_t>(num_); }
  const std::string& as_string() const { return str_; }
  // The literal text of a number as written in the document (e.g. "0.90").
  const std::string& raw_number() const { return str_; }

  size_t size() const { return is_array() ? arr_.size() : (is_object() ? members_.size() : 0); }
  const Value& at(size_t i) const { return arr_.at(i); }
  const std::vector<Value>& items() const { return arr_; }

  // Object access: nullptr when absent (or when this is not an object).
  const Value* find(const std::string& key) const {
    if (!is_object()) return nullptr;
    for (const auto& kv : members_)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool has(const std::string& key) const { return find(key) != nullptr; }
  const std::vector<std::pair<std::string, Value>>& members() const { return members_; }

  static Value parse(const char* text, size_t len) {
    Parser p{text, text + len};
    p.skip_ws();
    Value v = p.parse_value(0);
    p.skip_ws();
    if (p.cur != p.end) p.fail("trailing characters after JSON document");
    return v;
  }
  static Value parse(const std::string& text) { return parse(text.data(), text.size()); }

 private:
  Type type_ = Null;
  bool bool_ = false;
  bool integral_ = false;
  double num_ = 0.0;
  int64_t int_ = 0;
  std::string str_;
  std::vector<Value> arr_;
  std::vector<std::pair<std::string, Value>> members_;

  struct Parser {
    const char* cur;
    const char* end;
    const char* begin = cur;

    [[noreturn]] void fail(const std::string& what) const {
      throw std::runtime_error("JSON parse error at offset " + std::to_string(cur - begin) + ": " +
                               what);
    }
    void skip_ws() {
      while (cur != end && (*cur == ' ' || *cur == '\t' || *cur == '\n' || *cur == '\r')) ++cur;
    }
    bool consume(char c) {
      if (cur != end && *cur == c) {
        ++cur;
        return true;
      }
      return false;
    }
    void expect_word(const char* w) {
      for (const char* p = w; *p; ++p) {
        if (cur == end || *cur != *p) fail(std::string("expected '") + w + "'");
        ++cur;
      }
    }
    static void append_utf8(std::string& out, uint32_t cp) {
      if (cp < 0x80) {
        out.push_back(static_cast<char>(cp));
      } else if (cp < 0x800) {
        out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
      } else if (cp < 0x10000) {
        out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
        out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
      } else {
        out.push_back(static_cast<char>(0xF0 | (cp >> 18)));
        out.push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
        out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
      }
    }
    uint32_t parse_hex4() {
      uint32_t v = 0;
      for (int i = 0; i < 4; ++i) {
        if (cur == end) fail("truncated \\u escape");
        const char c = *cur++;
        v <<= 4;
        if (c >= '0' && c <= '9')
          v |= static_cast<uint32_t>(c - '0');
        else if (c >= 'a' && c <= 'f')
          v |= static_cast<uint32_t>(c - 'a' + 10);
        else if (c >= 'A' && c <= 'F')
          v |= static_cast<uint32_t>(c - 'A' + 10);
        else
          fail("bad hex digit in \\u escape");
      }
      return v;
    }
    std::string parse_string_body() {
      std::string out;
      while (true) {
        if (cur == end) fail("unterminated string");
        const char c = *cur++;
        if (c == '"') break;
        if (c != '\\') {
          out.push_back(c);
          continue;
        }
        if (cur == end) fail("unterminated escape");
        const char e = *cur++;
        switch (e) {
          case '"': out.push_back('"'); break;
          case '\\': out.push_back('\\'); break;
          case '/': out.push_back('/'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'n': out.push_back('\n'); break;
          case 'r': out.push_back('\r'); break;
          case 't': out.push_back('\t'); break;
          case 'u': {
            uint32_t cp = parse_hex4();
            if (cp >= 0xD800 && cp <= 0xDBFF && cur + 1 < end && cur[0] == '\\' && cur[1] == 'u') {
              cur += 2;
              const uint32_t lo = parse_hex4();
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            append_utf8(out, cp);
            break;
          }
          default: fail("unknown escape");
        }
      }
      return out;
    }
    Value parse_number() {
      const char* start = cur;
      if (cur != end && (*cur == '-' || *cur == '+')) ++cur;
      bool integral = true;
      while (cur != end) {
        const char c = *cur;
        if (c >= '0' && c <= '9') {
          ++cur;
        } else if (c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-') {
          integral = false;
          ++cur;
        } else {
          break;
        }
      }
      if (cur == start) fail("expected a value");
      Value v;
      v.type_ = Number;
      v.str_.assign(start, cur);
      char* endp = nullptr;
      v.num_ = std::strtod(v.str_.c_str(), &endp);
      if (endp == v.str_.c_str() || *endp != '\0') fail("malformed number '" + v.str_ + "'");
      v.integral_ = integral;
      if (integral) {
        // values above INT64_MAX (e.g. overflow_margin 2^64-1) saturate
        errno = 0;
        const long long ll = std::strtoll(v.str_.c_str(), nullptr, 10);
        v.int_ = static_cast<int64_t>(ll);
      }
      return v;
    }
    Value parse_value(int depth) {
      if (depth > 256) fail("nesting too deep");
      skip_ws();
      if (cur == end) fail("unexpected end of document");
      Value v;
      const char c = *cur;
      if (c == '{') {
        ++cur;
        v.type_ = Object;
        skip_ws();
        if (consume('}')) return v;
        while (true) {
          skip_ws();
          if (!consume('"')) fail("expected object key");
          std::string key = parse_string_body();
          skip_ws();
          if (!consume(':')) fail("expected ':'");
          Value member = parse_value(depth + 1);
          v.members_.emplace_back(std::move(key), std::move(member));
          skip_ws();
          if (consume(',')) continue;
          if (consume('}')) break;
          fail("expected ',' or '}'");
        }
      } else if (c == '[') {
        ++cur;
        v.type_ = Array;
        skip_ws();
        if (consume(']')) return v;
        while (true) {
          v.arr_.push_back(parse_value(depth + 1));
          skip_ws();
          if (consume(',')) continue;
          if (consume(']')) break;
          fail("expected ',' or ']'");
        }
      } else if (c == '"') {
        ++cur;
        v.type_ = String;
        v.str_ = parse_string_body();
      } else if (c == 't') {
        expect_word("true");
        v.type_ = Bool;
        v.bool_ = true;
      } else if (c == 'f') {
        expect_word("false");
        v.type_ = Bool;
        v.bool_ = false;
      } else if (c == 'n') {
        expect_word("null");
      } else {
        v = parse_number();
      }
      return v;
    }
  };
};

}  // namespace json
}  // namespace hpsx
