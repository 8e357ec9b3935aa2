// json.hpp — a small strict JSON reader for the `SimulationSpec` wire format.
// The reference round-trips this JSON through serde (crates/engine-wasm/src/lib.rs:203-230);
// here it is the C-ABI input of mcp_run_batch.  No third-party dependency.
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mcp {

struct JsonValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JsonValue> arr;
    std::map<std::string, JsonValue> obj;

    bool is_null() const { return kind == Null; }
    const JsonValue* get(const std::string& key) const {
        if (kind != Object) return nullptr;
        auto it = obj.find(key);
        return it == obj.end() ? nullptr : &it->second;
    }
};

class JsonParser {
  public:
    JsonParser(const char* s, size_t n) : p_(s), end_(s + n) {}
    JsonValue parse() {
        JsonValue v = value();
        ws();
        if (p_ != end_) fail("trailing characters");
        return v;
    }

  private:
    const char* p_;
    const char* end_;
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("spec JSON: ") + what); }
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
    }
    bool lit(const char* s) {
        size_t n = std::char_traits<char>::length(s);
        if ((size_t)(end_ - p_) >= n && std::char_traits<char>::compare(p_, s, n) == 0) {
            p_ += n;
            return true;
        }
        return false;
    }
    JsonValue value() {
        ws();
        if (p_ >= end_) fail("unexpected end");
        JsonValue v;
        char c = *p_;
        if (c == '{') {
            ++p_;
            v.kind = JsonValue::Object;
            ws();
            if (p_ < end_ && *p_ == '}') { ++p_; return v; }
            for (;;) {
                ws();
                if (p_ >= end_ || *p_ != '"') fail("expected object key");
                std::string k = string();
                ws();
                if (p_ >= end_ || *p_ != ':') fail("expected ':'");
                ++p_;
                v.obj[k] = value();
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == '}') { ++p_; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            ++p_;
            v.kind = JsonValue::Array;
            ws();
            if (p_ < end_ && *p_ == ']') { ++p_; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == ']') { ++p_; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.kind = JsonValue::String;
            v.str = string();
        } else if (lit("true")) {
            v.kind = JsonValue::Bool;
            v.b = true;
        } else if (lit("false")) {
            v.kind = JsonValue::Bool;
            v.b = false;
        } else if (lit("null")) {
            v.kind = JsonValue::Null;
        } else if (lit("NaN")) {
            v.kind = JsonValue::Number;
            v.num = std::nan("");
        } else if (lit("Infinity")) {
            v.kind = JsonValue::Number;
            v.num = INFINITY;
        } else if (lit("-Infinity")) {
            v.kind = JsonValue::Number;
            v.num = -INFINITY;
        } else {
            char* e = nullptr;
            std::string tmp(p_, (size_t)std::min<ptrdiff_t>(end_ - p_, 64));
            v.num = std::strtod(tmp.c_str(), &e);
            if (e == tmp.c_str()) fail("invalid token");
            p_ += (e - tmp.c_str());
            v.kind = JsonValue::Number;
        }
        return v;
    }
    std::string string() {
        ++p_;  // opening quote
        std::string out;
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\') {
                ++p_;
                if (p_ >= end_) fail("bad escape");
                switch (*p_) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (end_ - p_ < 5) fail("bad \\u escape");
                        unsigned code = (unsigned)std::strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
                        out += code < 0x80 ? (char)code : '?';
                        p_ += 4;
                        break;
                    }
                    default: out += *p_;
                }
                ++p_;
            } else {
                out += *p_++;
            }
        }
        if (p_ >= end_) fail("unterminated string");
        ++p_;
        return out;
    }
};

}  // namespace mcp
