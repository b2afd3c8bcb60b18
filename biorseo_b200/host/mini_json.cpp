#include "mini_json.h"

#include <cstdint>

namespace bsh {
namespace {

struct Cursor {
    const std::string& s;
    size_t             i = 0;
    std::string        err;
    explicit Cursor(const std::string& text) : s(text) {}

    bool fail(const char* what)
    {
        if (err.empty()) err = std::string(what) + " at byte " + std::to_string(i);
        return false;
    }
    void ws()
    {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++;
    }
    bool eat(char c)
    {
        ws();
        if (i < s.size() && s[i] == c) { i++; return true; }
        return false;
    }
    static void utf8(std::string& out, uint32_t cp)
    {
        if (cp < 0x80) out += char(cp);
        else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
        else { out += char(0xF0 | (cp >> 18)); out += char(0x80 | ((cp >> 12) & 0x3F)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
    }
    bool hex4(uint32_t& v)
    {
        if (i + 4 > s.size()) return fail("truncated \\u escape");
        v = 0;
        for (int k = 0; k < 4; k++) {
            char c = s[i++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
            else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
            else return fail("bad \\u escape");
        }
        return true;
    }
    bool string(std::string& out)
    {
        ws();
        if (i >= s.size() || s[i] != '"') return fail("expected string");
        i++;
        out.clear();
        while (i < s.size()) {
            unsigned char c = (unsigned char)s[i++];
            if (c == '"') return true;
            if (c < 0x20) return fail("control character in string");
            if (c != '\\') { out += char(c); continue; }
            if (i >= s.size()) break;
            char e = s[i++];
            switch (e) {
            case '"': out += '"'; break;
            case '\\': out += '\\'; break;
            case '/': out += '/'; break;
            case 'b': out += '\b'; break;
            case 'f': out += '\f'; break;
            case 'n': out += '\n'; break;
            case 'r': out += '\r'; break;
            case 't': out += '\t'; break;
            case 'u': {
                uint32_t cp = 0;
                if (!hex4(cp)) return false;
                if (cp >= 0xD800 && cp <= 0xDBFF) {
                    uint32_t lo = 0;
                    if (i + 2 > s.size() || s[i] != '\\' || s[i + 1] != 'u') return fail("lone surrogate");
                    i += 2;
                    if (!hex4(lo)) return false;
                    if (lo < 0xDC00 || lo > 0xDFFF) return fail("bad surrogate pair");
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                } else if (cp >= 0xDC00 && cp <= 0xDFFF) return fail("lone surrogate");
                utf8(out, cp);
                break;
            }
            default: return fail("bad escape");
            }
        }
        return fail("unterminated string");
    }
    bool literal(const char* word)
    {
        for (const char* p = word; *p; p++) {
            if (i >= s.size() || s[i] != *p) return fail("bad literal");
            i++;
        }
        return true;
    }
    bool number()
    {
        size_t start = i;
        if (i < s.size() && s[i] == '-') i++;
        while (i < s.size() && ((s[i] >= '0' && s[i] <= '9') || s[i] == '.' || s[i] == 'e' || s[i] == 'E' || s[i] == '+' || s[i] == '-')) i++;
        return i > start ? true : fail("expected value");
    }
    // Skips any value. When `captured` is non-null and the value is a string, stores it.
    bool value(std::string* captured, bool* was_string, int depth)
    {
        if (depth > 200) return fail("nesting too deep");
        ws();
        if (was_string) *was_string = false;
        if (i >= s.size()) return fail("unexpected end");
        char c = s[i];
        if (c == '"') {
            std::string tmp;
            if (!string(tmp)) return false;
            if (captured) *captured = tmp;
            if (was_string) *was_string = true;
            return true;
        }
        if (c == '{') {
            i++;
            if (eat('}')) return true;
            do {
                std::string k;
                if (!string(k)) return false;
                if (!eat(':')) return fail("expected ':'");
                if (!value(nullptr, nullptr, depth + 1)) return false;
            } while (eat(','));
            return eat('}') ? true : fail("expected '}'");
        }
        if (c == '[') {
            i++;
            if (eat(']')) return true;
            do {
                if (!value(nullptr, nullptr, depth + 1)) return false;
            } while (eat(','));
            return eat(']') ? true : fail("expected ']'");
        }
        if (c == 't') return literal("true");
        if (c == 'f') return literal("false");
        if (c == 'n') return literal("null");
        return number();
    }
    bool module(JsonModule& m)
    {
        ws();
        if (i >= s.size() || s[i] != '{') {   // not an object: remember that, skip it
            m = JsonModule();
            return value(nullptr, nullptr, 1);
        }
        i++;
        m = JsonModule();
        m.is_object = true;
        if (eat('}')) return true;
        do {
            std::string k, v;
            bool        is_str = false;
            if (!string(k)) return false;
            if (!eat(':')) return fail("expected ':'");
            if (!value(&v, &is_str, 2)) return false;
            if (k == "sequence") { m.has_sequence = is_str; m.sequence = is_str ? v : std::string(); }
            if (k == "struct2d") { m.has_struct2d = is_str; m.struct2d = is_str ? v : std::string(); }
        } while (eat(','));
        return eat('}') ? true : fail("expected '}'");
    }
};

}   // namespace

bool parse_module_json(const std::string& text, std::map<std::string, JsonModule>& out, std::string& err)
{
    Cursor c(text);
    out.clear();
    if (!c.eat('{')) { c.fail("top level must be an object"); err = c.err; return false; }
    if (!c.eat('}')) {
        do {
            std::string key;
            JsonModule  m;
            if (!c.string(key) || !c.eat(':') || !c.module(m)) {
                c.fail("malformed member");
                err = c.err;
                return false;
            }
            out[key] = m;
        } while (c.eat(','));
        if (!c.eat('}')) { c.fail("expected '}'"); err = c.err; return false; }
    }
    c.ws();
    if (c.i != text.size()) { c.fail("trailing characters"); err = c.err; return false; }
    return true;
}

}   // namespace bsh
