// Minimal JSON reader/writer for the three on-disk formats the hot path consumes (config JSON,
// network_models JSON) -- host side of libnbvh_b200.so.  Not a general-purpose library: numbers are
// doubles, arrays made only of numbers are stored flat (the weights file is one 10.8 M element array).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace nbvh {

struct Json {
    enum Type { Null, Bool, Number, String, Array, NumArray, Object } type = Null;
    bool b     = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<double> nums;
    std::vector<std::pair<std::string, Json>> obj;

    const Json *find(const char *key) const
    {
        if (type != Object)
            return nullptr;
        for (auto &kv : obj)
            if (kv.first == key)
                return &kv.second;
        return nullptr;
    }
    bool contains(const char *key) const { return find(key) != nullptr; }
    // json.value(key, default) semantics of the reference's config loader
    double value(const char *key, double def) const
    {
        const Json *j = find(key);
        return (j && j->type == Number) ? j->num : def;
    }
    bool value(const char *key, bool def) const
    {
        const Json *j = find(key);
        return (j && j->type == Bool) ? j->b : def;
    }
    std::string value(const char *key, const std::string &def) const
    {
        const Json *j = find(key);
        return (j && j->type == String) ? j->str : def;
    }
    const Json &object(const char *key) const
    {
        static const Json empty_object = [] {
            Json j;
            j.type = Object;
            return j;
        }();
        const Json *j = find(key);
        return (j && j->type == Object) ? *j : empty_object;
    }
    size_t size() const { return type == NumArray ? nums.size() : arr.size(); }
    double at_num(size_t i) const
    {
        if (type == NumArray)
            return nums.at(i);
        if (type == Array && arr.at(i).type == Number)
            return arr[i].num;
        throw std::runtime_error("json: element is not a number");
    }
};

class JsonParser {
public:
    JsonParser(const char *text, size_t len) : p(text), end(text + len) {}
    Json parse()
    {
        Json j = value();
        ws();
        if (p != end)
            fail("trailing characters");
        return j;
    }

private:
    const char *p, *end;
    [[noreturn]] void fail(const char *msg) { throw std::runtime_error(std::string("json: ") + msg); }
    void ws()
    {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r'))
            ++p;
    }
    Json value()
    {
        ws();
        if (p >= end)
            fail("unexpected end");
        switch (*p) {
        case '{':
            return object();
        case '[':
            return array();
        case '"': {
            Json j;
            j.type = Json::String;
            j.str  = string();
            return j;
        }
        case 't':
        case 'f':
        case 'n':
            return literal();
        default:
            return number();
        }
    }
    Json literal()
    {
        Json j;
        if (end - p >= 4 && !strncmp(p, "true", 4)) {
            j.type = Json::Bool, j.b = true, p += 4;
        } else if (end - p >= 5 && !strncmp(p, "false", 5)) {
            j.type = Json::Bool, j.b = false, p += 5;
        } else if (end - p >= 4 && !strncmp(p, "null", 4)) {
            j.type = Json::Null, p += 4;
        } else
            fail("bad literal");
        return j;
    }
    Json number()
    {
        // fast path for small non-negative integers (the weights byte array)
        const char *s = p;
        if (*p >= '0' && *p <= '9') {
            uint64_t v = 0;
            while (p < end && *p >= '0' && *p <= '9')
                v = v * 10 + uint64_t(*p++ - '0');
            if (p >= end || (*p != '.' && *p != 'e' && *p != 'E')) {
                Json j;
                j.type = Json::Number;
                j.num  = double(v);
                return j;
            }
            p = s;
        }
        char *e  = nullptr;
        double v = strtod(p, &e);
        if (e == p)
            fail("bad number");
        p = e;
        Json j;
        j.type = Json::Number;
        j.num  = v;
        return j;
    }
    std::string string()
    {
        ++p;
        std::string s;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end)
                    fail("bad escape");
                switch (*p) {
                case 'n': s += '\n'; break;
                case 't': s += '\t'; break;
                case 'r': s += '\r'; break;
                case 'b': s += '\b'; break;
                case 'f': s += '\f'; break;
                case 'u': {
                    if (end - p < 5)
                        fail("bad unicode escape");
                    unsigned code = (unsigned)strtoul(std::string(p + 1, 4).c_str(), nullptr, 16);
                    s += code < 0x80 ? char(code) : '?';
                    p += 4;
                    break;
                }
                default: s += *p;
                }
                ++p;
            } else
                s += *p++;
        }
        if (p >= end)
            fail("unterminated string");
        ++p;
        return s;
    }
    Json array()
    {
        ++p;
        Json j;
        j.type         = Json::NumArray;
        bool all_nums  = true;
        ws();
        if (p < end && *p == ']') {
            ++p;
            j.type = Json::Array;
            return j;
        }
        for (;;) {
            Json v = value();
            if (all_nums && v.type == Json::Number) {
                j.nums.push_back(v.num);
            } else {
                if (all_nums) { // demote what we have so far
                    all_nums = false;
                    j.type   = Json::Array;
                    for (double d : j.nums) {
                        Json n;
                        n.type = Json::Number;
                        n.num  = d;
                        j.arr.push_back(n);
                    }
                    j.nums.clear();
                }
                j.arr.push_back(std::move(v));
            }
            ws();
            if (p >= end)
                fail("unterminated array");
            if (*p == ',') {
                ++p;
                continue;
            }
            if (*p == ']') {
                ++p;
                break;
            }
            fail("expected , or ]");
        }
        return j;
    }
    Json object()
    {
        ++p;
        Json j;
        j.type = Json::Object;
        ws();
        if (p < end && *p == '}') {
            ++p;
            return j;
        }
        for (;;) {
            ws();
            if (p >= end || *p != '"')
                fail("expected key");
            std::string k = string();
            ws();
            if (p >= end || *p != ':')
                fail("expected :");
            ++p;
            j.obj.emplace_back(std::move(k), value());
            ws();
            if (p >= end)
                fail("unterminated object");
            if (*p == ',') {
                ++p;
                continue;
            }
            if (*p == '}') {
                ++p;
                break;
            }
            fail("expected , or }");
        }
        return j;
    }
};

inline Json json_parse(const std::string &text) { return JsonParser(text.data(), text.size()).parse(); }

} // namespace nbvh
