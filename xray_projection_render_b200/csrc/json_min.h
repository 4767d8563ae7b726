// json_min.h -- minimal JSON reader for scene / deformation descriptions.
// Numbers are parsed with strtod (correctly rounded), so a float64 written by Go's
// encoding/json or Python's json round-trips bit-exactly.
#pragma once
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace xr {

struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;

    const JValue* get(const char* key) const {
        if (kind != Obj) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_num() const { return kind == Num; }
};

class JParser {
public:
    explicit JParser(const char* s) : p_(s) {}
    bool parse(JValue& out, std::string& err) {
        if (!value(out, 0)) {
            err = err_.empty() ? "invalid JSON" : err_;
            return false;
        }
        ws();
        if (*p_) {
            err = "trailing characters after JSON value";
            return false;
        }
        return true;
    }

private:
    const char* p_;
    std::string err_;
    void ws() {
        while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r') ++p_;
    }
    bool fail(const char* m) {
        if (err_.empty()) err_ = m;
        return false;
    }
    bool value(JValue& v, int depth) {
        if (depth > 64) return fail("JSON nesting too deep");
        ws();
        switch (*p_) {
            case '{': return object(v, depth);
            case '[': return array(v, depth);
            case '"': v.kind = JValue::Str; return string(v.str);
            case 't':
                if (!strncmp(p_, "true", 4)) { p_ += 4; v.kind = JValue::Bool; v.b = true; return true; }
                return fail("bad literal");
            case 'f':
                if (!strncmp(p_, "false", 5)) { p_ += 5; v.kind = JValue::Bool; v.b = false; return true; }
                return fail("bad literal");
            case 'n':
                if (!strncmp(p_, "null", 4)) { p_ += 4; v.kind = JValue::Null; return true; }
                return fail("bad literal");
            default: return number(v);
        }
    }
    bool number(JValue& v) {
        const char* s = p_;
        if (*p_ == '-' || *p_ == '+') ++p_;
        if (!strncmp(p_, "Infinity", 8) || !strncmp(p_, "NaN", 3)) return fail("non-finite number");
        char* e = nullptr;
        double d = strtod(s, &e);
        if (e == s) return fail("bad number");
        p_ = e;
        v.kind = JValue::Num;
        v.num = d;
        return true;
    }
    bool string(std::string& out) {
        ++p_;  // opening quote
        out.clear();
        while (*p_ && *p_ != '"') {
            if (*p_ == '\\') {
                ++p_;
                switch (*p_) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        unsigned cp = 0;
                        for (int i = 1; i <= 4; ++i) {
                            char c = p_[i];
                            if (!c) return fail("bad \\u escape");
                            cp = cp * 16 + (c <= '9' ? c - '0' : (c | 32) - 'a' + 10);
                        }
                        p_ += 4;
                        if (cp < 0x80) out += char(cp);
                        else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
                        else { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    case 0: return fail("unterminated string");
                    default: out += *p_;
                }
                ++p_;
            } else {
                out += *p_++;
            }
        }
        if (*p_ != '"') return fail("unterminated string");
        ++p_;
        return true;
    }
    bool array(JValue& v, int depth) {
        ++p_;
        v.kind = JValue::Arr;
        ws();
        if (*p_ == ']') { ++p_; return true; }
        while (true) {
            v.arr.emplace_back();
            if (!value(v.arr.back(), depth + 1)) return false;
            ws();
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == ']') { ++p_; return true; }
            return fail("expected , or ] in array");
        }
    }
    bool object(JValue& v, int depth) {
        ++p_;
        v.kind = JValue::Obj;
        ws();
        if (*p_ == '}') { ++p_; return true; }
        while (true) {
            ws();
            if (*p_ != '"') return fail("expected string key");
            std::string k;
            if (!string(k)) return false;
            ws();
            if (*p_ != ':') return fail("expected : after key");
            ++p_;
            v.obj.emplace_back(k, JValue());
            if (!value(v.obj.back().second, depth + 1)) return false;
            ws();
            if (*p_ == ',') { ++p_; continue; }
            if (*p_ == '}') { ++p_; return true; }
            return fail("expected , or } in object");
        }
    }
};

}  // namespace xr
