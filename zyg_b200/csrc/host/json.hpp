// Minimal JSON reader for the strings the C API takes (su_material_create, su_integrators_create, ...)
// and for the take / scene subset. Objects keep insertion order like std.json.ObjectMap, which the
// reference iterates when it applies material parameters (material_provider.zig:131-161).
#pragma once

#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace zyg::json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;

    bool                                       boolean = false;
    double                                     number  = 0.0;
    std::string                                string;
    std::vector<Value>                         array;
    std::vector<std::pair<std::string, Value>> object;

    const Value* get(const char* key) const {
        if (Object != kind) return nullptr;
        for (const auto& kv : object) {
            if (kv.first == key) return &kv.second;
        }
        return nullptr;
    }
    bool isNumber() const { return Number == kind; }
};

class Parser {
  public:
    explicit Parser(const char* text) : p_(text) {}

    bool parse(Value& out) {
        skip();
        if (!value(out)) return false;
        skip();
        return '\0' == *p_;
    }

  private:
    const char* p_;

    void skip() {
        while (' ' == *p_ || '\n' == *p_ || '\t' == *p_ || '\r' == *p_) ++p_;
    }
    bool literal(const char* word) {
        const size_t n = std::strlen(word);
        if (0 != std::strncmp(p_, word, n)) return false;
        p_ += n;
        return true;
    }
    bool string(std::string& out) {
        if ('"' != *p_) return false;
        ++p_;
        out.clear();
        while ('"' != *p_) {
            if ('\0' == *p_) return false;
            if ('\\' == *p_) {
                ++p_;
                switch (*p_) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break;
                    case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'u': {  // BMP code points only; enough for ASCII-range escapes
                        char hex[5] = {0, 0, 0, 0, 0};
                        for (int i = 0; i < 4; ++i) {
                            if ('\0' == p_[1 + i]) return false;
                            hex[i] = p_[1 + i];
                        }
                        out.push_back(char(std::strtoul(hex, nullptr, 16)));
                        p_ += 4;
                        break;
                    }
                    case '\0': return false;
                    default: out.push_back(*p_); break;
                }
                ++p_;
            } else {
                out.push_back(*p_++);
            }
        }
        ++p_;
        return true;
    }
    bool value(Value& out) {
        skip();
        switch (*p_) {
            case '{': {
                ++p_;
                out.kind = Value::Object;
                skip();
                if ('}' == *p_) {
                    ++p_;
                    return true;
                }
                for (;;) {
                    skip();
                    std::string key;
                    if (!string(key)) return false;
                    skip();
                    if (':' != *p_) return false;
                    ++p_;
                    Value v;
                    if (!value(v)) return false;
                    out.object.emplace_back(std::move(key), std::move(v));
                    skip();
                    if (',' == *p_) {
                        ++p_;
                        continue;
                    }
                    if ('}' == *p_) {
                        ++p_;
                        return true;
                    }
                    return false;
                }
            }
            case '[': {
                ++p_;
                out.kind = Value::Array;
                skip();
                if (']' == *p_) {
                    ++p_;
                    return true;
                }
                for (;;) {
                    Value v;
                    if (!value(v)) return false;
                    out.array.push_back(std::move(v));
                    skip();
                    if (',' == *p_) {
                        ++p_;
                        continue;
                    }
                    if (']' == *p_) {
                        ++p_;
                        return true;
                    }
                    return false;
                }
            }
            case '"': out.kind = Value::String; return string(out.string);
            case 't':
                out.kind    = Value::Bool;
                out.boolean = true;
                return literal("true");
            case 'f':
                out.kind    = Value::Bool;
                out.boolean = false;
                return literal("false");
            case 'n': out.kind = Value::Null; return literal("null");
            default: {
                char*        end = nullptr;
                const double d   = std::strtod(p_, &end);
                if (end == p_) return false;
                p_         = end;
                out.kind   = Value::Number;
                out.number = d;
                return true;
            }
        }
    }
};

// src/base/json.zig readers
inline float readFloat(const Value& v) { return float(v.number); }
inline float readFloatMember(const Value& v, const char* name, float def) {
    const Value* m = v.get(name);
    return m && m->isNumber() ? float(m->number) : def;
}
inline uint32_t readUIntMember(const Value& v, const char* name, uint32_t def) {
    const Value* m = v.get(name);
    return m && m->isNumber() ? uint32_t(m->number) : def;
}
inline bool readBoolMember(const Value& v, const char* name, bool def) {
    const Value* m = v.get(name);
    return m && Value::Bool == m->kind ? m->boolean : def;
}

}  // namespace zyg::json
