// Minimal JSON reader for HF config.json (objects, arrays, numbers, strings, true/false/null).
#pragma once

#include <cstdlib>
#include <cstring>
#include <map>
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
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("config json: ") + m); }
    void ws() { while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r') ++p_; }
    Value value() {
        ws();
        Value v;
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
    std::string str() {
        std::string out;
        ++p_;
        while (*p_ && *p_ != '"') {
            if (*p_ == '\\') {
                ++p_;
                switch (*p_) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'u': p_ += 4; out += '?'; break;
                    default: out += *p_;
                }
                ++p_;
            } else out += *p_++;
        }
        if (*p_ != '"') fail("unterminated string");
        ++p_;
        return out;
    }
};

}  // namespace cbjson
