// rt_json.h — minimal JSON reader for the scene files (the reference uses rapidjson; only the subset its loader
// touches is needed: objects, arrays, strings, numbers, booleans).  Numbers remember whether they were written as
// integers because the reference's loader distinguishes IsInt / IsFloat (Demo/SceneLoader.cpp:101-150).
#pragma once

#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace rt {
namespace json {

struct Value
{
    enum class Type { Null, Bool, Number, String, Array, Object } type = Type::Null;
    bool boolean = false;
    double number = 0.0;
    bool isInteger = false;      // written without '.', 'e' or 'E'
    std::string string;
    std::vector<Value> array;
    std::vector<std::pair<std::string, Value>> object;

    bool IsObject() const { return type == Type::Object; }
    bool IsArray() const { return type == Type::Array; }
    bool IsString() const { return type == Type::String; }
    bool IsBool() const { return type == Type::Bool; }
    bool IsNumber() const { return type == Type::Number; }
    // rapidjson: IsInt() for integer literals that fit int32, IsFloat() only for non-integer literals in float range
    bool IsInt() const { return type == Type::Number && isInteger && number >= -2147483648.0 && number <= 2147483647.0; }
    bool IsFloat() const { return type == Type::Number && !isInteger && number >= -3.4028234e38 && number <= 3.4028234e38; }
    float GetFloat() const { return static_cast<float>(number); }
    int GetInt() const { return static_cast<int>(number); }
    size_t Size() const { return array.size(); }
    const Value& operator[](size_t i) const { return array[i]; }
    bool HasMember(const char* name) const { return Find(name) != nullptr; }
    const Value* Find(const char* name) const
    {
        for (const auto& kv : object) if (kv.first == name) return &kv.second;
        return nullptr;
    }
    const Value& operator[](const char* name) const { static const Value null; const Value* v = Find(name); return v ? *v : null; }
};

class Parser
{
public:
    bool Parse(const std::string& text, Value& out, std::string& err)
    {
        p = text.c_str(); end = p + text.size();
        if (!ParseValue(out)) { err = error.empty() ? "syntax error" : error; return false; }
        SkipWs();
        if (p != end) { err = "trailing characters after the JSON document"; return false; }
        return true;
    }

private:
    const char* p = nullptr; const char* end = nullptr;
    std::string error;

    void SkipWs() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool Fail(const char* m) { if (error.empty()) error = m; return false; }

    int depth = 0;
    struct DepthGuard { int& d; explicit DepthGuard(int& x) : d(x) { ++d; } ~DepthGuard() { --d; } };

    bool ParseValue(Value& v)
    {
        const DepthGuard guard(depth);
        if (depth > 256) return Fail("nesting deeper than 256 levels");     // recursion bound for hostile input
        SkipWs();
        if (p >= end) return Fail("unexpected end of input");
        switch (*p)
        {
        case '{': return ParseObject(v);
        case '[': return ParseArray(v);
        case '"': v.type = Value::Type::String; return ParseString(v.string);
        case 't': if (end - p >= 4 && !strncmp(p, "true", 4)) { p += 4; v.type = Value::Type::Bool; v.boolean = true; return true; } return Fail("invalid literal");
        case 'f': if (end - p >= 5 && !strncmp(p, "false", 5)) { p += 5; v.type = Value::Type::Bool; v.boolean = false; return true; } return Fail("invalid literal");
        case 'n': if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; v.type = Value::Type::Null; return true; } return Fail("invalid literal");
        default: return ParseNumber(v);
        }
    }

    bool ParseNumber(Value& v)
    {
        const char* s = p;
        if (p < end && (*p == '-' || *p == '+')) ++p;
        bool digits = false, integer = true;
        while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+'))
        {
            if (*p >= '0' && *p <= '9') digits = true;
            if (*p == '.' || *p == 'e' || *p == 'E') integer = false;
            ++p;
        }
        if (!digits) return Fail("invalid number");
        const std::string tok(s, p);
        v.type = Value::Type::Number;
        v.number = strtod(tok.c_str(), nullptr);
        v.isInteger = integer;
        return true;
    }

    bool ParseString(std::string& out)
    {
        ++p;    // opening quote
        out.clear();
        while (p < end && *p != '"')
        {
            if (*p == '\\')
            {
                if (++p >= end) return Fail("bad escape");
                switch (*p)
                {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': { if (end - p < 5) return Fail("bad \\u escape"); const unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16); out += (char)(cp & 0x7F); p += 4; break; }
                default: out += *p; break;
                }
                ++p;
            }
            else out += *p++;
        }
        if (p >= end) return Fail("unterminated string");
        ++p;
        return true;
    }

    bool ParseArray(Value& v)
    {
        v.type = Value::Type::Array;
        ++p;
        SkipWs();
        if (p < end && *p == ']') { ++p; return true; }
        for (;;)
        {
            v.array.emplace_back();
            if (!ParseValue(v.array.back())) return false;
            SkipWs();
            if (p < end && *p == ',') { ++p; continue; }
            if (p < end && *p == ']') { ++p; return true; }
            return Fail("expected ',' or ']'");
        }
    }

    bool ParseObject(Value& v)
    {
        v.type = Value::Type::Object;
        ++p;
        SkipWs();
        if (p < end && *p == '}') { ++p; return true; }
        for (;;)
        {
            SkipWs();
            if (p >= end || *p != '"') return Fail("expected a member name");
            std::string key;
            if (!ParseString(key)) return false;
            SkipWs();
            if (p >= end || *p != ':') return Fail("expected ':'");
            ++p;
            v.object.emplace_back(key, Value());
            if (!ParseValue(v.object.back().second)) return false;
            SkipWs();
            if (p < end && *p == ',') { ++p; continue; }
            if (p < end && *p == '}') { ++p; return true; }
            return Fail("expected ',' or '}'");
        }
    }
};

} // namespace json
} // namespace rt
