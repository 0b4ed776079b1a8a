// Minimal JSON reader for HF config.json and safetensors headers (objects, arrays, numbers, strings, true/false/null).
// The input is untrusted (file headers): every read is bounded by the terminating NUL, nesting is limited, \uXXXX escapes
// (surrogate pairs included) are decoded to UTF-8.
#pragma once

#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace cbjson {

struct Value {
    enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> obj;

    bool has(const std::string& k) const { return type == OBJ && obj.count(k) && obj.at(k).type != NUL; }
    const Value& at(const std::string& k) const {
        if (!has(k)) throw std::runtime_error("config: missing key '" + k + "'");
        return obj.at(k);
    }
    double number(const std::string& k, double dflt) const { return has(k) ? obj.at(k).num : dflt; }
    long long integer(const std::string& k, long long dflt) const { return has(k) ? (long long)obj.at(k).num : dflt; }
    long long integer(const std::string& k) const { return (long long)at(k).num; }
    bool boolean(const std::string& k, bool dflt) const { return has(k) ? obj.at(k).b : dflt; }
    std::string string(const std::string& k, const std::string& dflt) const { return has(k) ? obj.at(k).str : dflt; }
};

class Parser {
public:
    explicit Parser(const char* s) : p_(s) {}
    Value parse() {
        Value v = value();
        ws();
        if (*p_) fail("trailing characters");
        return v;
    }

private:
    const char* p_;
    int depth_ = 0;
    static constexpr int kMaxDepth = 64;
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("config json: ") + m); }
    void ws() { while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r') ++p_; }
    struct Nest {
        int& d;
        explicit Nest(int& dd) : d(dd) { ++d; }
        ~Nest() { --d; }
    };
    Value value() {
        ws();
        Value v;
        Nest nest(depth_);
        if (depth_ > kMaxDepth) fail("nesting too deep");
        if (*p_ == '{') {
            v.type = Value::OBJ;
            ++p_; ws();
            if (*p_ == '}') { ++p_; return v; }
            for (;;) {
                ws();
                if (*p_ != '"') fail("expected string key");
                std::string k = str();
                ws();
                if (*p_ != ':') fail("expected ':'");
                ++p_;
                v.obj[k] = value();
                ws();
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == '}') { ++p_; break; }
                fail("expected ',' or '}'");
            }
        } else if (*p_ == '[') {
            v.type = Value::ARR;
            ++p_; ws();
            if (*p_ == ']') { ++p_; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (*p_ == ',') { ++p_; continue; }
                if (*p_ == ']') { ++p_; break; }
                fail("expected ',' or ']'");
            }
        } else if (*p_ == '"') {
            v.type = Value::STR;
            v.str = str();
        } else if (!std::strncmp(p_, "true", 4)) { v.type = Value::BOOL; v.b = true; p_ += 4; }
        else if (!std::strncmp(p_, "false", 5)) { v.type = Value::BOOL; v.b = false; p_ += 5; }
        else if (!std::strncmp(p_, "null", 4)) { v.type = Value::NUL; p_ += 4; }
        else {
            char* end = nullptr;
            v.num = std::strtod(p_, &end);
            if (end == p_) fail("unexpected token");
            v.type = Value::NUM;
            p_ = end;
        }
        return v;
    }
    int hex4() {               // four hex digits at p_ (each checked: a NUL ends the buffer)
        int v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = p_[i];
            int d;
            if (c >= '0' && c <= '9') d = c - '0';
            else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
            else fail("bad \\u escape");
            v = v * 16 + d;
        }
        p_ += 4;
        return v;
    }
    static void utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    std::string str() {
        std::string out;
        ++p_;
        while (*p_ && *p_ != '"') {
            if (*p_ != '\\') { out += *p_++; continue; }
            ++p_;
            const char c = *p_;
            if (c == 0) fail("unterminated string");
            ++p_;
            switch (c) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': {
                    unsigned cp = (unsigned)hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && p_[0] == '\\' && p_[1] == 'u') {      // surrogate pair
                        p_ += 2;
                        const unsigned lo = (unsigned)hex4();
                        if (lo < 0xDC00 || lo > 0xDFFF) fail("bad surrogate pair");
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(out, cp);
                    break;
                }
                default: out += c;          // \" \\ \/
            }
        }
        if (*p_ != '"') fail("unterminated string");
        ++p_;
        return out;
    }
};

}  // namespace cbjson
