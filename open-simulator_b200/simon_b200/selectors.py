"""Object-level label / node-selector / toleration matching (host side).

  labels.Requirement.Matches            vendor/k8s.io/apimachinery/pkg/labels/selector.go:200-244
  metav1.LabelSelectorAsSelector        vendor/k8s.io/apimachinery/pkg/apis/meta/v1/helpers.go
  nodeaffinity term matching            vendor/k8s.io/component-helpers/scheduling/corev1/nodeaffinity/nodeaffinity.go:82-257
  PodMatchesNodeSelectorAndAffinityTerms vendor/k8s.io/kubernetes/pkg/scheduler/framework/plugins/helper/node_affinity.go:27-71
  Toleration.ToleratesTaint             vendor/k8s.io/api/core/v1/toleration.go:37-56
"""
from __future__ import annotations

from typing import Dict, List, Optional

from .objects import Obj


class SelectorError(ValueError):
    pass


# ---- canonical requirement lists: [(key, op, (values...))] with op in In/NotIn/Exists/DoesNotExist/Gt/Lt ----

def label_selector_requirements(sel: Optional[Obj]):
    """metav1.LabelSelectorAsSelector. Returns None for a nil selector (matches nothing),
    [] for the empty selector (matches everything), else the requirement list."""
    if sel is None:
        return None
    reqs = []
    for k, v in sorted((sel.get("matchLabels") or {}).items()):
        reqs.append((k, "In", (str(v),)))
    for e in sel.get("matchExpressions") or []:
        op = e.get("operator")
        vals = tuple(sorted(str(x) for x in (e.get("values") or [])))
        if op in ("In", "NotIn"):
            if not vals:
                raise SelectorError("for 'in', 'notin' operators, values set can't be empty")
        elif op in ("Exists", "DoesNotExist"):
            if vals:
                raise SelectorError("values set must be empty for exists and does not exist")
        else:
            raise SelectorError(f"{op!r} is not a valid pod selector operator")
        reqs.append((e.get("key"), op, vals))
    return reqs


def requirement_matches(req, labels: Dict[str, str]) -> bool:
    key, op, vals = req
    if op == "In":
        return key in labels and labels[key] in vals
    if op == "NotIn":
        return key not in labels or labels[key] not in vals
    if op == "Exists":
        return key in labels
    if op == "DoesNotExist":
        return key not in labels
    if op in ("Gt", "Lt"):
        if key not in labels:
            return False
        try:
            lv = _parse_int64(labels[key])
        except ValueError:
            return False
        if len(vals) != 1:
            return False
        try:
            rv = _parse_int64(vals[0])
        except ValueError:
            return False
        return lv > rv if op == "Gt" else lv < rv
    return False


def _parse_int64(s: str) -> int:
    """strconv.ParseInt(s, 10, 64): optional sign, decimal digits only, range-checked."""
    t = s[1:] if s[:1] in "+-" else s
    if not t or not t.isascii() or not t.isdigit():
        raise ValueError(s)
    v = int(s)
    if not (-(1 << 63) <= v < (1 << 63)):
        raise ValueError(s)
    return v


def requirements_match(reqs, labels: Dict[str, str]) -> bool:
    if reqs is None:
        return False
    return all(requirement_matches(r, labels) for r in reqs)


def label_selector_matches(sel: Optional[Obj], labels: Dict[str, str]) -> bool:
    return requirements_match(label_selector_requirements(sel), labels)


# ---- node selector terms ----

def node_selector_term_requirements(term: Obj):
    """newNodeSelectorTerm: returns (label_reqs|None, field_reqs|None, ok). An invalid term never matches."""
    lab = None
    fld = None
    exprs = term.get("matchExpressions") or []
    fields = term.get("matchFields") or []
    if exprs:
        lab = []
        for e in exprs:
            op = e.get("operator")
            vals = tuple(str(x) for x in (e.get("values") or []))
            if op in ("In", "NotIn"):
                if not vals:
                    return None, None, False
            elif op in ("Exists", "DoesNotExist"):
                if vals:
                    return None, None, False
            elif op in ("Gt", "Lt"):
                if len(vals) != 1:
                    return None, None, False
                try:
                    _parse_int64(vals[0])
                except ValueError:
                    return None, None, False
            else:
                return None, None, False
            lab.append((e.get("key"), op, vals))
    if fields:
        fld = []
        for e in fields:
            op = e.get("operator")
            vals = [str(x) for x in (e.get("values") or [])]
            if op not in ("In", "NotIn") or len(vals) != 1:
                return None, None, False
            fld.append((e.get("key"), op, vals[0]))
    return lab, fld, True


def is_empty_node_selector_term(term: Obj) -> bool:
    return not (term.get("matchExpressions") or []) and not (term.get("matchFields") or [])


def node_selector_term_matches(term: Obj, node: Obj) -> bool:
    lab, fld, ok = node_selector_term_requirements(term)
    if not ok:
        return False
    labels = (node.get("metadata") or {}).get("labels") or {}
    if lab is not None and not all(requirement_matches(r, labels) for r in lab):
        return False
    name = (node.get("metadata") or {}).get("name", "")
    if fld is not None and len(name) > 0:
        for key, op, val in fld:
            have = name if key == "metadata.name" else None
            # fields.Set only holds metadata.name; other keys read as "" (fields.Set.Get)
            have = have if have is not None else ""
            if op == "In" and have != val:
                return False
            if op == "NotIn" and have == val:
                return False
    return True


def node_matches_node_selector(node: Obj, node_selector: Obj) -> bool:
    """LazyErrorNodeSelector.Match: terms ORed; empty terms are skipped; no terms -> no match."""
    for term in node_selector.get("nodeSelectorTerms") or []:
        if is_empty_node_selector_term(term):
            continue
        if node_selector_term_matches(term, node):
            return True
    return False


def pod_matches_node_selector_and_affinity(spec: Obj, node: Obj) -> bool:
    labels = (node.get("metadata") or {}).get("labels") or {}
    ns = spec.get("nodeSelector") or {}
    if len(ns) > 0:
        for k, v in ns.items():
            if labels.get(k) != str(v) or k not in labels:
                return False
    aff = spec.get("affinity")
    if aff is None:
        return True
    na = aff.get("nodeAffinity")
    if na is None:
        return True
    req = na.get("requiredDuringSchedulingIgnoredDuringExecution")
    if req is not None and not node_matches_node_selector(node, req):
        return False
    return True


# ---- taints ----

def toleration_tolerates_taint(tol: Obj, taint: Obj) -> bool:
    eff = tol.get("effect") or ""
    if len(eff) > 0 and eff != (taint.get("effect") or ""):
        return False
    key = tol.get("key") or ""
    if len(key) > 0 and key != (taint.get("key") or ""):
        return False
    op = tol.get("operator") or ""
    if op in ("", "Equal"):
        return (tol.get("value") or "") == (taint.get("value") or "")
    if op == "Exists":
        return True
    return False


def tolerations_tolerate_taint(tols: List[Obj], taint: Obj) -> bool:
    return any(toleration_tolerates_taint(t, taint) for t in tols)
