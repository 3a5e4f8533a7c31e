// Host side (native): API objects as JSON-shaped values.  The Go caller marshals its typed objects with encoding/json
// (every k8s API type carries json tags), the library parses them here; key order of objects is preserved (class keys and
// interning orders below are first-appearance orders, as in the Python host side this file mirrors: simon_b200/objects.py).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace sh {

struct Error : std::runtime_error {
    int code;      // simon_status
    explicit Error(const std::string &m, int c = -1) : std::runtime_error(m), code(c) {}
};

struct J {
    enum T : uint8_t { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false;
    std::string s;                                  // Str: the string; Num: the literal as written
    std::vector<J> a;
    std::vector<std::pair<std::string, J>> o;

    J() = default;
    static J str(std::string v) { J j; j.t = Str; j.s = std::move(v); return j; }
    static J num(long long v) { J j; j.t = Num; j.s = std::to_string(v); return j; }
    static J boolean(bool v) { J j; j.t = Bool; j.b = v; return j; }
    static J obj() { J j; j.t = Obj; return j; }
    static J arr() { J j; j.t = Arr; return j; }

    bool is_obj() const { return t == Obj; }
    bool is_arr() const { return t == Arr; }
    bool is_str() const { return t == Str; }
    bool is_null() const { return t == Null; }

    const J *get(const char *k) const {
        if (t != Obj) return nullptr;
        for (auto &kv : o) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const J *get(const std::string &k) const { return get(k.c_str()); }
    J *getm(const char *k) {
        if (t != Obj) return nullptr;
        for (auto &kv : o) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    bool has(const char *k) const { return get(k) != nullptr; }
    // dict[k] = v (existing key keeps its position, like a Python dict)
    J &set(const std::string &k, J v) {
        if (t != Obj) { *this = obj(); }
        for (auto &kv : o) if (kv.first == k) { kv.second = std::move(v); return kv.second; }
        o.emplace_back(k, std::move(v));
        return o.back().second;
    }
    void erase(const char *k) {
        if (t != Obj) return;
        for (size_t i = 0; i < o.size(); i++) if (o[i].first == k) { o.erase(o.begin() + (long)i); return; }
    }
    // Python truthiness
    bool truthy() const {
        switch (t) {
            case Null: return false;
            case Bool: return b;
            case Num: return strtod(s.c_str(), nullptr) != 0.0;
            case Str: return !s.empty();
            case Arr: return !a.empty();
            case Obj: return !o.empty();
        }
        return false;
    }
    // str(value) of a scalar as the Python host side sees it after json.loads
    std::string text() const {
        switch (t) {
            case Str: return s;
            case Num: return s;
            case Bool: return b ? "True" : "False";
            case Null: return "None";
            default: return "";
        }
    }
    long long as_int(long long dflt = 0) const {
        if (t == Num) return (long long)strtod(s.c_str(), nullptr) == strtoll(s.c_str(), nullptr, 10) ? strtoll(s.c_str(), nullptr, 10)
                                                                                                         : (long long)strtod(s.c_str(), nullptr);
        if (t == Str) { char *e = nullptr; long long v = strtoll(s.c_str(), &e, 10); return (e && *e == 0 && !s.empty()) ? v : dflt; }
        if (t == Bool) return b ? 1 : 0;
        return dflt;
    }
    size_t size() const { return t == Arr ? a.size() : t == Obj ? o.size() : 0; }
};

static const J kNullJ;
static const J kEmptyObj = J::obj();
static const J kEmptyArr = J::arr();

// `x.get(k) or {}` / `x.get(k) or []`
inline const J &or_obj(const J *p) { return (p && p->t == J::Obj && !p->o.empty()) ? *p : kEmptyObj; }
inline const J &or_arr(const J *p) { return (p && p->t == J::Arr && !p->a.empty()) ? *p : kEmptyArr; }
inline const J &field_obj(const J &x, const char *k) { return or_obj(x.get(k)); }
inline const J &field_arr(const J &x, const char *k) { return or_arr(x.get(k)); }
// `x.get(k) or ""` for strings
inline std::string field_str(const J &x, const char *k) {
    const J *p = x.get(k);
    return (p && p->truthy()) ? p->text() : std::string();
}
inline bool present(const J *p) { return p && !p->is_null(); }     // `x.get(k) is not None`

// ---- canonical text (class keys, memo keys): key order preserved unless sort_keys ----
inline void dump_str(std::string &out, const std::string &s) {
    out.push_back('"');
    for (unsigned char c : s) {
        if (c == '"' || c == '\\') { out.push_back('\\'); out.push_back((char)c); }
        else if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
        else out.push_back((char)c);
    }
    out.push_back('"');
}

inline void dump(std::string &out, const J &j, bool sort_keys = false) {
    switch (j.t) {
        case J::Null: out += "null"; break;
        case J::Bool: out += j.b ? "true" : "false"; break;
        case J::Num: out += j.s; break;
        case J::Str: dump_str(out, j.s); break;
        case J::Arr:
            out.push_back('[');
            for (size_t i = 0; i < j.a.size(); i++) { if (i) out.push_back(','); dump(out, j.a[i], sort_keys); }
            out.push_back(']');
            break;
        case J::Obj: {
            out.push_back('{');
            if (!sort_keys) {
                for (size_t i = 0; i < j.o.size(); i++) {
                    if (i) out.push_back(',');
                    dump_str(out, j.o[i].first); out.push_back(':'); dump(out, j.o[i].second, false);
                }
            } else {
                std::vector<size_t> idx(j.o.size());
                for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
                for (size_t i = 1; i < idx.size(); i++)       // insertion sort: objects are small
                    for (size_t k = i; k > 0 && j.o[idx[k]].first < j.o[idx[k - 1]].first; k--) std::swap(idx[k], idx[k - 1]);
                for (size_t i = 0; i < idx.size(); i++) {
                    if (i) out.push_back(',');
                    dump_str(out, j.o[idx[i]].first); out.push_back(':'); dump(out, j.o[idx[i]].second, true);
                }
            }
            out.push_back('}');
            break;
        }
    }
}
inline void dump_opt(std::string &out, const J *p, bool sort_keys = false) { if (p) dump(out, *p, sort_keys); else out += "null"; }

// ---- parser (RFC 8259; encoding/json never emits duplicate keys: were there one, lookups see the first) ----
struct Parser {
    const char *p, *e;
    explicit Parser(const char *s, size_t n) : p(s), e(s + n) {}
    [[noreturn]] void fail(const char *m) const { throw Error(std::string("request JSON: ") + m); }
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    static void utf8(std::string &out, uint32_t c) {
        if (c < 0x80) out.push_back((char)c);
        else if (c < 0x800) { out.push_back((char)(0xC0 | (c >> 6))); out.push_back((char)(0x80 | (c & 0x3F))); }
        else if (c < 0x10000) { out.push_back((char)(0xE0 | (c >> 12))); out.push_back((char)(0x80 | ((c >> 6) & 0x3F))); out.push_back((char)(0x80 | (c & 0x3F))); }
        else { out.push_back((char)(0xF0 | (c >> 18))); out.push_back((char)(0x80 | ((c >> 12) & 0x3F))); out.push_back((char)(0x80 | ((c >> 6) & 0x3F))); out.push_back((char)(0x80 | (c & 0x3F))); }
    }
    uint32_t hex4() {
        if (e - p < 4) fail("truncated \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string string() {
        if (p >= e || *p != '"') fail("expected string");
        p++;
        std::string out;
        while (true) {
            if (p >= e) fail("unterminated string");
            const char *q = p;
            while (q < e && *q != '"' && *q != '\\') q++;
            out.append(p, q);
            p = q;
            if (p >= e) fail("unterminated string");
            if (*p == '"') { p++; return out; }
            p++;
            if (p >= e) fail("bad escape");
            char c = *p++;
            switch (c) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t c1 = hex4();
                    if (c1 >= 0xD800 && c1 < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2;
                        uint32_t c2 = hex4();
                        if (c2 >= 0xDC00 && c2 < 0xE000) c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
                        else { utf8(out, c1); c1 = c2; }
                    }
                    utf8(out, c1);
                    break;
                }
                default: fail("bad escape");
            }
        }
    }
    // ---- large arrays (cluster.Nodes, cluster.Pods: tens of thousands of objects) are parsed by several threads: one sequential
    //      scan finds the element boundaries (strings and nesting only), then each thread parses a contiguous run of elements ----
    void skip_string() {
        p++;
        while (true) {
            if (p >= e) fail("unterminated string");
            if (*p == '"') { p++; return; }
            if (*p == '\\') { p++; if (p >= e) fail("bad escape"); }
            p++;
        }
    }
    void skip_value() {
        ws();
        if (p >= e) fail("unexpected end");
        if (*p == '"') { skip_string(); return; }
        if (*p == '{' || *p == '[') {
            int nest = 0;
            while (true) {
                if (p >= e) fail("unexpected end");
                const char ch = *p;
                if (ch == '"') { skip_string(); continue; }
                if (ch == '{' || ch == '[') nest++;
                else if (ch == '}' || ch == ']') { nest--; if (nest == 0) { p++; return; } }
                p++;
            }
        }
        while (p < e && *p != ',' && *p != ']' && *p != '}' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;     // literal / number
    }
    // p stands on the first element of an array; true: `j.a` holds every element and p stands after the closing bracket
    bool parallel_array(J &j, int depth) {
        const char *start = p;
        std::vector<std::pair<const char *, const char *>> spans;
        while (true) {
            ws();
            const char *b = p;
            skip_value();
            spans.emplace_back(b, p);
            ws();
            if (p < e && *p == ',') { p++; continue; }
            if (p < e && *p == ']') { p++; break; }
            fail("expected ',' or ']'");
        }
        unsigned hw = std::thread::hardware_concurrency();
        const size_t nt = std::min<size_t>(8, std::min<size_t>(hw ? hw : 1, spans.size() / 64));
        if (nt < 2) { p = start; return false; }
        const char *after = p;
        j.a.resize(spans.size());
        std::vector<std::string> errors(nt);
        std::vector<std::thread> th;
        for (size_t t = 0; t < nt; t++)
            th.emplace_back([&, t] {
                const size_t lo = spans.size() * t / nt, hi = spans.size() * (t + 1) / nt;
                try {
                    for (size_t i = lo; i < hi; i++) {
                        Parser sub(spans[i].first, (size_t)(spans[i].second - spans[i].first));
                        j.a[i] = sub.value(depth + 1);
                        sub.ws();
                        if (sub.p != sub.e) sub.fail("unexpected character");
                    }
                } catch (const std::exception &ex) { errors[t] = ex.what(); if (errors[t].empty()) errors[t] = "request JSON: error"; }
            });
        for (auto &x : th) x.join();
        for (auto &m : errors) if (!m.empty()) throw Error(m);
        p = after;
        return true;
    }
    J value(int depth = 0) {
        if (depth > 200) fail("nesting too deep");
        ws();
        if (p >= e) fail("unexpected end");
        J j;
        char c = *p;
        if (c == '{') {
            p++;
            j.t = J::Obj;
            ws();
            if (p < e && *p == '}') { p++; return j; }
            while (true) {
                ws();
                std::string k = string();
                ws();
                if (p >= e || *p != ':') fail("expected ':'");
                p++;
                if (j.o.empty()) j.o.reserve(8);
                j.o.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; return j; }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            p++;
            j.t = J::Arr;
            ws();
            if (p < e && *p == ']') { p++; return j; }
            if (depth <= 3 && e - p > (1 << 18) && parallel_array(j, depth)) return j;
            while (true) {
                j.a.push_back(value(depth + 1));
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == ']') { p++; return j; }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') { j.t = J::Str; j.s = string(); return j; }
        if (c == 't' && e - p >= 4 && !memcmp(p, "true", 4)) { p += 4; j.t = J::Bool; j.b = true; return j; }
        if (c == 'f' && e - p >= 5 && !memcmp(p, "false", 5)) { p += 5; j.t = J::Bool; j.b = false; return j; }
        if (c == 'n' && e - p >= 4 && !memcmp(p, "null", 4)) { p += 4; return j; }
        if (c == '-' || (c >= '0' && c <= '9')) {
            const char *q = p;
            if (*q == '-') q++;
            while (q < e && ((*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '+' || *q == '-')) q++;
            j.t = J::Num;
            j.s.assign(p, q);
            p = q;
            return j;
        }
        fail("unexpected character");
    }
};

inline J parse_json(const char *s, size_t n) {
    Parser ps(s, n);
    J v = ps.value();
    ps.ws();
    if (ps.p != ps.e) ps.fail("trailing characters");
    return v;
}

// ---- accessors shared by the modules (simon_b200/objects.py) ----
inline const J &meta_of(const J &o) { return field_obj(o, "metadata"); }
inline std::string name_of(const J &o) { return field_str(meta_of(o), "name"); }
inline std::string namespace_of(const J &o) { return field_str(meta_of(o), "namespace"); }
inline const J &labels_of(const J &o) { return field_obj(meta_of(o), "labels"); }
inline const J &annotations_of(const J &o) { return field_obj(meta_of(o), "annotations"); }
inline const J &spec_of(const J &o) { return field_obj(o, "spec"); }

static const char *const ANNO_NODE_LOCAL_STORAGE = "simon/node-local-storage";   // pkg/type/const.go:21
static const char *const ANNO_POD_LOCAL_STORAGE = "simon/pod-local-storage";
static const char *const ANNO_WORKLOAD_KIND = "simon/workload-kind";
static const char *const ANNO_WORKLOAD_NAME = "simon/workload-name";
static const char *const ANNO_WORKLOAD_NAMESPACE = "simon/workload-namespace";
static const char *const LABEL_NEW_NODE = "simon/new-node";
static const char *const LABEL_APP_NAME = "simon/app-name";
static const char *const LABEL_HOSTNAME = "kubernetes.io/hostname";
static const char *const LABEL_ZONE = "topology.kubernetes.io/zone";
static const char *const LABEL_REGION = "topology.kubernetes.io/region";
static const char *const LABEL_ZONE_BETA = "failure-domain.beta.kubernetes.io/zone";
static const char *const LABEL_REGION_BETA = "failure-domain.beta.kubernetes.io/region";

}  // namespace sh
