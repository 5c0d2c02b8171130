// json_min.h -- a small recursive-descent JSON reader, enough for OpenMVG's cereal
// sfm_data.json (objects, arrays, numbers, strings, true/false/null), and a writer that puts a
// parsed tree back with every number it did not touch spelled as it was read.  Header only.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace hulo {
namespace json {

struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    const Value *get(const char *key) const {
        if (type != Object) return nullptr;
        for (const auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    // follows a path of keys; nullptr when any is missing
    const Value *path(std::initializer_list<const char *> keys) const {
        const Value *v = this;
        for (const char *k : keys) {
            if (!v) return nullptr;
            v = v->get(k);
        }
        return v;
    }
    double number(double dflt = 0.0) const { return type == Number ? num : dflt; }

    Value *find(const char *key) { return const_cast<Value *>(static_cast<const Value *>(this)->get(key)); }
    // a number set by the program (written with 17 significant digits)
    void set_number(double v) { type = Number; num = v; str.clear(); }
};

inline void dump_string(const std::string &in, std::string &out) {
    out += '"';
    for (char c : in) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\t': out += "\\t"; break;
            case '\r': out += "\\r"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            default:
                if ((unsigned char)c < 0x20) {
                    char buf[8];
                    snprintf(buf, sizeof buf, "\\u%04x", (unsigned)(unsigned char)c);
                    out += buf;
                } else {
                    out += c;       // UTF-8 bytes pass through
                }
        }
    }
    out += '"';
}

// cereal's layout: 4 spaces per level, one member per line
inline void dump(const Value &v, std::string &out, int indent = 0) {
    const std::string pad((size_t)(indent + 1) * 4, ' '), pad_close((size_t)indent * 4, ' ');
    switch (v.type) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Number:
            if (!v.str.empty()) {
                out += v.str;                  // as read
            } else {
                char buf[40];
                snprintf(buf, sizeof buf, "%.17g", v.num);
                out += buf;
            }
            break;
        case Value::String: dump_string(v.str, out); break;
        case Value::Array:
            if (v.arr.empty()) { out += "[]"; break; }
            out += "[\n";
            for (size_t i = 0; i < v.arr.size(); ++i) {
                out += pad;
                dump(v.arr[i], out, indent + 1);
                out += i + 1 < v.arr.size() ? ",\n" : "\n";
            }
            out += pad_close + "]";
            break;
        case Value::Object:
            if (v.obj.empty()) { out += "{}"; break; }
            out += "{\n";
            for (size_t i = 0; i < v.obj.size(); ++i) {
                out += pad;
                dump_string(v.obj[i].first, out);
                out += ": ";
                dump(v.obj[i].second, out, indent + 1);
                out += i + 1 < v.obj.size() ? ",\n" : "\n";
            }
            out += pad_close + "}";
            break;
    }
}

class Parser {
public:
    explicit Parser(const std::string &text) : s_(text.c_str()), end_(text.c_str() + text.size()) {}
    bool parse(Value &out) {
        skip();
        if (!value(out)) return false;
        skip();
        return true;
    }

private:
    const char *s_, *end_;
    void skip() {
        while (s_ < end_ && (*s_ == ' ' || *s_ == '\n' || *s_ == '\t' || *s_ == '\r')) ++s_;
    }
    bool literal(const char *w) {
        const size_t n = strlen(w);
        if ((size_t)(end_ - s_) < n || strncmp(s_, w, n) != 0) return false;
        s_ += n;
        return true;
    }
    bool string(std::string &out) {
        if (s_ >= end_ || *s_ != '"') return false;
        ++s_;
        out.clear();
        while (s_ < end_ && *s_ != '"') {
            if (*s_ == '\\' && s_ + 1 < end_) {
                ++s_;
                switch (*s_) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {   // \uXXXX -> UTF-8 (surrogate pairs joined)
                        auto hex4 = [&](const char *p, unsigned &v) {
                            v = 0;
                            for (int k = 0; k < 4; ++k) {
                                const char c = p[k];
                                v <<= 4;
                                if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
                                else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
                                else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
                                else return false;
                            }
                            return true;
                        };
                        unsigned cp = 0;
                        if (end_ - s_ < 5 || !hex4(s_ + 1, cp)) return false;
                        s_ += 4;
                        if (cp >= 0xD800 && cp < 0xDC00 && end_ - s_ >= 7 && s_[1] == '\\' && s_[2] == 'u') {
                            unsigned lo = 0;
                            if (hex4(s_ + 3, lo) && lo >= 0xDC00 && lo < 0xE000) {
                                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                                s_ += 6;
                            }
                        }
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += *s_;
                }
                ++s_;
            } else {
                out += *s_++;
            }
        }
        if (s_ >= end_) return false;
        ++s_;
        return true;
    }
    bool value(Value &v) {
        skip();
        if (s_ >= end_) return false;
        const char c = *s_;
        if (c == '{') {
            v.type = Value::Object;
            ++s_;
            skip();
            if (s_ < end_ && *s_ == '}') { ++s_; return true; }
            for (;;) {
                skip();
                std::string key;
                if (!string(key)) return false;
                skip();
                if (s_ >= end_ || *s_ != ':') return false;
                ++s_;
                v.obj.emplace_back(std::move(key), Value());
                if (!value(v.obj.back().second)) return false;
                skip();
                if (s_ < end_ && *s_ == ',') { ++s_; continue; }
                if (s_ < end_ && *s_ == '}') { ++s_; return true; }
                return false;
            }
        }
        if (c == '[') {
            v.type = Value::Array;
            ++s_;
            skip();
            if (s_ < end_ && *s_ == ']') { ++s_; return true; }
            for (;;) {
                v.arr.emplace_back();
                if (!value(v.arr.back())) return false;
                skip();
                if (s_ < end_ && *s_ == ',') { ++s_; continue; }
                if (s_ < end_ && *s_ == ']') { ++s_; return true; }
                return false;
            }
        }
        if (c == '"') { v.type = Value::String; return string(v.str); }
        if (literal("true")) { v.type = Value::Bool; v.b = true; return true; }
        if (literal("false")) { v.type = Value::Bool; v.b = false; return true; }
        if (literal("null")) { v.type = Value::Null; return true; }
        char *e = nullptr;
        v.num = strtod(s_, &e);
        if (e == s_) return false;
        v.type = Value::Number;
        v.str.assign(s_, (size_t)(e - s_));    // the spelling, so an untouched number is written back as read
        s_ = e;
        return true;
    }
};

}  // namespace json
}  // namespace hulo
