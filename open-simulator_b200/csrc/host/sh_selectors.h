// Object-level label / node-selector / toleration matching (native host side; mirrors simon_b200/selectors.py).
//   labels.Requirement.Matches             vendor/k8s.io/apimachinery/pkg/labels/selector.go:200-244
//   metav1.LabelSelectorAsSelector         vendor/k8s.io/apimachinery/pkg/apis/meta/v1/helpers.go
//   nodeaffinity term matching             vendor/k8s.io/component-helpers/scheduling/corev1/nodeaffinity/nodeaffinity.go:82-257
//   PodMatchesNodeSelectorAndAffinityTerms vendor/k8s.io/kubernetes/pkg/scheduler/framework/plugins/helper/node_affinity.go:27-71
//   Toleration.ToleratesTaint              vendor/k8s.io/api/core/v1/toleration.go:37-56
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "sh_json.h"

namespace sh {

struct SelectorError : Error { explicit SelectorError(const std::string &m) : Error(m) {} };

enum ReqOp : uint8_t { OP_IN, OP_NOTIN, OP_EXISTS, OP_DOESNOTEXIST, OP_GT, OP_LT, OP_BAD };

inline ReqOp op_of(const std::string &s) {
    if (s == "In") return OP_IN;
    if (s == "NotIn") return OP_NOTIN;
    if (s == "Exists") return OP_EXISTS;
    if (s == "DoesNotExist") return OP_DOESNOTEXIST;
    if (s == "Gt") return OP_GT;
    if (s == "Lt") return OP_LT;
    return OP_BAD;
}
inline const char *op_name(ReqOp o) {
    static const char *n[] = {"In", "NotIn", "Exists", "DoesNotExist", "Gt", "Lt", "?"};
    return n[o];
}

struct Req {
    std::string key;
    ReqOp op;
    std::vector<std::string> vals;
    bool operator==(const Req &r) const { return key == r.key && op == r.op && vals == r.vals; }
};
// a selector: `none` = the nil selector (matches nothing); else the AND of reqs (empty = matches everything)
struct ReqList {
    bool none = false;
    std::vector<Req> reqs;
    std::string key() const {
        if (none) return "~";
        std::string k;
        for (auto &r : reqs) {
            k += r.key; k.push_back('\x01'); k += op_name(r.op);
            for (auto &v : r.vals) { k.push_back('\x02'); k += v; }
            k.push_back('\x03');
        }
        return k;
    }
};

// string-keyed label maps are read straight off the JSON objects
inline const J *label_get(const J &labels, const std::string &k) { return labels.get(k); }
inline std::string label_text(const J &v) { return v.is_null() ? std::string() : v.text(); }

// strconv.ParseInt(s, 10, 64)
inline bool parse_int64(const std::string &s, long long &out) {
    size_t k = (!s.empty() && (s[0] == '+' || s[0] == '-')) ? 1 : 0;
    if (k >= s.size()) return false;
    for (size_t i = k; i < s.size(); i++) if (s[i] < '0' || s[i] > '9') return false;
    __int128 v = 0;
    for (size_t i = k; i < s.size(); i++) {
        v = v * 10 + (s[i] - '0');
        if (v > ((__int128)1 << 63)) return false;
    }
    if (s[0] == '-') v = -v;
    if (v < -((__int128)1 << 63) || v >= ((__int128)1 << 63)) return false;
    out = (long long)v;
    return true;
}

// metav1.LabelSelectorAsSelector
inline ReqList label_selector_requirements(const J *sel) {
    ReqList out;
    if (!sel || sel->is_null()) { out.none = true; return out; }
    const J &ml = field_obj(*sel, "matchLabels");
    std::vector<std::pair<std::string, std::string>> pairs;
    for (auto &kv : ml.o) pairs.emplace_back(kv.first, kv.second.text());
    std::sort(pairs.begin(), pairs.end());
    for (auto &p : pairs) out.reqs.push_back(Req{p.first, OP_IN, {p.second}});
    for (auto &e : field_arr(*sel, "matchExpressions").a) {
        const J *opj = e.get("operator");
        ReqOp op = opj && opj->is_str() ? op_of(opj->s) : OP_BAD;
        std::vector<std::string> vals;
        for (auto &x : field_arr(e, "values").a) vals.push_back(x.text());
        std::sort(vals.begin(), vals.end());
        if (op == OP_IN || op == OP_NOTIN) {
            if (vals.empty()) throw SelectorError("for 'in', 'notin' operators, values set can't be empty");
        } else if (op == OP_EXISTS || op == OP_DOESNOTEXIST) {
            if (!vals.empty()) throw SelectorError("values set must be empty for exists and does not exist");
        } else {
            throw SelectorError("'" + (opj ? opj->text() : std::string("None")) + "' is not a valid pod selector operator");
        }
        const J *kj = e.get("key");
        out.reqs.push_back(Req{kj ? kj->text() : std::string(), op, vals});
    }
    return out;
}

inline bool requirement_matches(const Req &r, const J &labels) {
    const J *lv = labels.get(r.key);
    switch (r.op) {
        case OP_IN: return lv && std::find(r.vals.begin(), r.vals.end(), label_text(*lv)) != r.vals.end();
        case OP_NOTIN: return !lv || std::find(r.vals.begin(), r.vals.end(), label_text(*lv)) == r.vals.end();
        case OP_EXISTS: return lv != nullptr;
        case OP_DOESNOTEXIST: return lv == nullptr;
        case OP_GT:
        case OP_LT: {
            if (!lv) return false;
            long long a, b;
            if (!parse_int64(label_text(*lv), a)) return false;
            if (r.vals.size() != 1 || !parse_int64(r.vals[0], b)) return false;
            return r.op == OP_GT ? a > b : a < b;
        }
        default: return false;
    }
}
inline bool requirements_match(const ReqList &rl, const J &labels) {
    if (rl.none) return false;
    for (auto &r : rl.reqs) if (!requirement_matches(r, labels)) return false;
    return true;
}

// ---- node selector terms (newNodeSelectorTerm) ----
struct FieldReq { std::string key; ReqOp op; std::string val; };
struct TermReqs {
    bool ok = true;
    bool has_lab = false, has_fld = false;
    std::vector<Req> lab;
    std::vector<FieldReq> fld;
};

inline TermReqs node_selector_term_requirements(const J &term) {
    TermReqs t;
    const J &exprs = field_arr(term, "matchExpressions");
    const J &fields = field_arr(term, "matchFields");
    if (!exprs.a.empty()) {
        t.has_lab = true;
        for (auto &e : exprs.a) {
            const J *opj = e.get("operator");
            ReqOp op = opj && opj->is_str() ? op_of(opj->s) : OP_BAD;
            std::vector<std::string> vals;
            for (auto &x : field_arr(e, "values").a) vals.push_back(x.text());
            if (op == OP_IN || op == OP_NOTIN) { if (vals.empty()) { t.ok = false; return t; } }
            else if (op == OP_EXISTS || op == OP_DOESNOTEXIST) { if (!vals.empty()) { t.ok = false; return t; } }
            else if (op == OP_GT || op == OP_LT) {
                long long v;
                if (vals.size() != 1 || !parse_int64(vals[0], v)) { t.ok = false; return t; }
            } else { t.ok = false; return t; }
            const J *kj = e.get("key");
            t.lab.push_back(Req{kj ? kj->text() : std::string(), op, vals});
        }
    }
    if (!fields.a.empty()) {
        t.has_fld = true;
        for (auto &e : fields.a) {
            const J *opj = e.get("operator");
            ReqOp op = opj && opj->is_str() ? op_of(opj->s) : OP_BAD;
            std::vector<std::string> vals;
            for (auto &x : field_arr(e, "values").a) vals.push_back(x.text());
            if ((op != OP_IN && op != OP_NOTIN) || vals.size() != 1) { t.ok = false; return t; }
            const J *kj = e.get("key");
            t.fld.push_back(FieldReq{kj ? kj->text() : std::string(), op, vals[0]});
        }
    }
    return t;
}

inline bool is_empty_node_selector_term(const J &term) {
    return field_arr(term, "matchExpressions").a.empty() && field_arr(term, "matchFields").a.empty();
}

inline bool node_selector_term_matches(const J &term, const J &node) {
    TermReqs t = node_selector_term_requirements(term);
    if (!t.ok) return false;
    const J &labels = labels_of(node);
    if (t.has_lab) for (auto &r : t.lab) if (!requirement_matches(r, labels)) return false;
    const J *nm = meta_of(node).get("name");
    std::string name = nm ? (nm->is_null() ? std::string("None") : nm->text()) : std::string();
    if (t.has_fld && !name.empty()) {
        for (auto &f : t.fld) {
            std::string have = f.key == "metadata.name" ? name : std::string();      // fields.Set only holds metadata.name
            if (f.op == OP_IN && have != f.val) return false;
            if (f.op == OP_NOTIN && have == f.val) return false;
        }
    }
    return true;
}

// LazyErrorNodeSelector.Match: terms ORed; empty terms are skipped; no terms -> no match
inline bool node_matches_node_selector(const J &node, const J &node_selector) {
    for (auto &term : field_arr(node_selector, "nodeSelectorTerms").a) {
        if (is_empty_node_selector_term(term)) continue;
        if (node_selector_term_matches(term, node)) return true;
    }
    return false;
}

inline bool pod_matches_node_selector_and_affinity(const J &spec, const J &node) {
    const J &labels = labels_of(node);
    const J &ns = field_obj(spec, "nodeSelector");
    for (auto &kv : ns.o) {
        const J *lv = labels.get(kv.first);
        // labels.get(k) != str(v) or k not in labels
        if (!lv || lv->t != J::Str || lv->s != kv.second.text()) return false;
    }
    const J *aff = spec.get("affinity");
    if (!present(aff)) return true;
    const J *na = aff->get("nodeAffinity");
    if (!present(na)) return true;
    const J *req = na->get("requiredDuringSchedulingIgnoredDuringExecution");
    if (present(req) && !node_matches_node_selector(node, *req)) return false;
    return true;
}

// ---- taints ----
struct Taint { std::string key, value, effect; };

inline bool toleration_tolerates_taint(const J &tol, const Taint &t) {
    std::string eff = field_str(tol, "effect");
    if (!eff.empty() && eff != t.effect) return false;
    std::string key = field_str(tol, "key");
    if (!key.empty() && key != t.key) return false;
    std::string op = field_str(tol, "operator");
    if (op.empty() || op == "Equal") return field_str(tol, "value") == t.value;
    if (op == "Exists") return true;
    return false;
}
inline bool tolerations_tolerate_taint(const J &tols, const Taint &t) {
    for (auto &x : tols.a) if (toleration_tolerates_taint(x, t)) return true;
    return false;
}
inline Taint taint_of(const J &t) { return Taint{field_str(t, "key"), field_str(t, "value"), field_str(t, "effect")}; }

}  // namespace sh
