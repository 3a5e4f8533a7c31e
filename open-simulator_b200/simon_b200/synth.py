"""Seeded synthetic clusters: Kubernetes objects as dicts, so that synthetic inputs travel the
same path as YAML inputs (workload expansion -> snapshot compiler -> engine).

PRNG = SplitMix64, seed 0x51514D4F4E00 + config number.
  C2: N nodes (8 shapes), 16 zones, hostname label; Deployments x replicas with cpu/mem requests,
      5 % request nothing; one Service per workload (activates the system default spreading).
  C3: + regions, taints, 32 label keys x <=8 values, tolerations, nodeSelector, required node affinity,
      required hostname anti-affinity, preferred zone affinity, explicit topologySpreadConstraints,
      10 % of nodes pre-filled with ~20 running pods.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

from .objects import AppResource, ResourceTypes

MASK = (1 << 64) - 1
SEED_BASE = 0x51514D4F4E00


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.next() % n

    def chance(self, pct: int) -> bool:
        return self.below(100) < pct

    def pick(self, seq):
        return seq[self.below(len(seq))]


CPU_SHAPES = [8, 16, 32, 64, 96, 128, 192, 256]


def _node(name: str, cores: int, zone: str, region: str, eph_gi: int, labels=None, taints=None, pods=110) -> dict:
    lab = {"kubernetes.io/hostname": name, "topology.kubernetes.io/zone": zone}
    if region:
        lab["topology.kubernetes.io/region"] = region
    if labels:
        lab.update(labels)
    n = {"apiVersion": "v1", "kind": "Node", "metadata": {"name": name, "labels": lab},
         "spec": {},
         "status": {"allocatable": {"cpu": str(cores), "memory": f"{4 * cores}Gi", "ephemeral-storage": f"{eph_gi}Gi",
                                    "pods": str(pods)},
                    "capacity": {"cpu": str(cores), "memory": f"{4 * cores}Gi", "ephemeral-storage": f"{eph_gi}Gi",
                                 "pods": str(pods)}}}
    if taints:
        n["spec"]["taints"] = taints
    return n


def _log_uniform(rng: SplitMix64, lo: int, hi: int, step: int) -> int:
    """log-uniform integer in [lo, hi], rounded to a multiple of step."""
    import math
    u = rng.below(1 << 20) / float(1 << 20)
    v = math.exp(math.log(lo) + u * (math.log(hi) - math.log(lo)))
    q = max(step, int(round(v / step)) * step)
    return min(q, hi)


def _deployment(name: str, ns: str, replicas: int, labels: Dict[str, str], cpu_m: int, mem_mi: int, spec_extra=None) -> dict:
    c = {"name": "c", "image": f"registry.local/{name}:v1"}
    if cpu_m or mem_mi:
        req = {}
        if cpu_m:
            req["cpu"] = f"{cpu_m}m"
        if mem_mi:
            req["memory"] = f"{mem_mi}Mi"
        c["resources"] = {"requests": req}
    spec = {"containers": [c]}
    if spec_extra:
        spec.update(spec_extra)
    return {"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": name, "namespace": ns},
            "spec": {"replicas": replicas, "selector": {"matchLabels": dict(labels)},
                     "template": {"metadata": {"labels": dict(labels)}, "spec": spec}}}


def _service(name: str, ns: str, sel: Dict[str, str]) -> dict:
    return {"apiVersion": "v1", "kind": "Service", "metadata": {"name": name, "namespace": ns},
            "spec": {"selector": dict(sel), "ports": [{"port": 80}]}}


def make_c2(n_nodes: int = 1000, n_workloads: int = 100, replicas: int = 100, seed_no: int = 2,
            n_apps: int = 1) -> Tuple[ResourceTypes, List[AppResource]]:
    rng = SplitMix64(SEED_BASE + seed_no)
    cluster = ResourceTypes()
    for i in range(n_nodes):
        cores = rng.pick(CPU_SHAPES)
        eph = rng.pick([100, 200, 500, 1000, 2048])
        cluster.Nodes.append(_node(f"node-{i:05d}", cores, f"zone-{i % 16:02d}", "", eph))
    apps = [AppResource(f"app-{a}", ResourceTypes()) for a in range(n_apps)]
    for w in range(n_workloads):
        name = f"wl-{w:04d}"
        ns = "default"
        labels = {"app": name}
        if rng.chance(5):
            cpu, mem = 0, 0
        else:
            cpu = _log_uniform(rng, 100, 8000, 50)
            mem = _log_uniform(rng, 128, 32768, 64)
        apps[w % n_apps].Resource.Deployments.append(_deployment(name, ns, replicas, labels, cpu, mem))
        cluster.Services.append(_service(name, ns, labels))
    return cluster, apps


def make_c3(n_nodes: int = 10000, n_workloads: int = 1000, replicas: int = 100, seed_no: int = 3,
            n_apps: int = 10, prefill_pct: int = 10, prefill_pods: int = 20) -> Tuple[ResourceTypes, List[AppResource]]:
    rng = SplitMix64(SEED_BASE + seed_no)
    cluster = ResourceTypes()
    hard_taints = [{"key": f"dedicated-{i}", "value": f"team-{i}", "effect": "NoSchedule"} for i in range(8)]
    soft_taints = [{"key": f"degraded-{i}", "value": "true", "effect": "PreferNoSchedule"} for i in range(4)]
    label_keys = [f"feature-{k:02d}" for k in range(32)]
    for i in range(n_nodes):
        cores = rng.pick(CPU_SHAPES)
        eph = rng.pick([100, 200, 500, 1000, 2048])
        region = f"region-{i % 3}"
        zone = f"{region}-zone-{(i // 3) % 16:02d}"
        labels = {}
        for k in label_keys:
            if rng.chance(50):
                labels[k] = f"v{rng.below(8)}"
        taints = []
        if rng.chance(10):
            taints.append(dict(rng.pick(hard_taints)))
        if rng.chance(5):
            taints.append(dict(rng.pick(soft_taints)))
        cluster.Nodes.append(_node(f"node-{i:05d}", cores, zone, region, eph, labels, taints or None))
    # running pods on ~prefill_pct % of the nodes (pre-bound: spec.nodeName set)
    for i in range(n_nodes):
        if not rng.chance(prefill_pct):
            continue
        for j in range(prefill_pods):
            cluster.Pods.append({"apiVersion": "v1", "kind": "Pod",
                                 "metadata": {"name": f"running-{i:05d}-{j:02d}", "namespace": "kube-system",
                                              "labels": {"app": f"sys-{j % 4}"}},
                                 "spec": {"nodeName": f"node-{i:05d}",
                                          "containers": [{"name": "c", "image": "registry.local/sys:v1",
                                                          "resources": {"requests": {"cpu": "100m", "memory": "128Mi"}}}]}})
    apps = [AppResource(f"app-{a}", ResourceTypes()) for a in range(n_apps)]
    names = [f"wl-{w:04d}" for w in range(n_workloads)]
    for w in range(n_workloads):
        name = names[w]
        ns = f"ns-{w % 8}"
        labels = {"app": name, "tier": rng.pick(["web", "db", "cache", "batch"])}
        if rng.chance(5):
            cpu, mem = 0, 0
        else:
            cpu = _log_uniform(rng, 100, 8000, 50)
            mem = _log_uniform(rng, 128, 32768, 64)
        extra: dict = {}
        aff: dict = {}
        if rng.chance(30):
            t = rng.pick(hard_taints)
            tols = [{"key": t["key"], "operator": "Equal", "value": t["value"], "effect": "NoSchedule"}]
            if rng.chance(50):
                s = rng.pick(soft_taints)
                tols.append({"key": s["key"], "operator": "Exists"})
            extra["tolerations"] = tols
        if rng.chance(20):
            sel = {}
            for _ in range(1 + rng.below(2)):
                sel[rng.pick(label_keys)] = f"v{rng.below(8)}"
            extra["nodeSelector"] = sel
        if rng.chance(10):
            exprs = []
            kind = rng.below(3)
            k = rng.pick(label_keys)
            if kind == 0:
                exprs.append({"key": k, "operator": "In", "values": [f"v{rng.below(8)}", f"v{rng.below(8)}", f"v{rng.below(8)}"]})
            elif kind == 1:
                exprs.append({"key": k, "operator": "NotIn", "values": [f"v{rng.below(8)}"]})
            else:
                exprs.append({"key": k, "operator": "Exists"})
            aff["nodeAffinity"] = {"requiredDuringSchedulingIgnoredDuringExecution":
                                   {"nodeSelectorTerms": [{"matchExpressions": exprs}]}}
            if rng.chance(50):
                aff["nodeAffinity"]["preferredDuringSchedulingIgnoredDuringExecution"] = [
                    {"weight": 1 + rng.below(100),
                     "preference": {"matchExpressions": [{"key": rng.pick(label_keys), "operator": "In",
                                                          "values": [f"v{rng.below(8)}"]}]}}]
        if rng.chance(20):
            aff["podAntiAffinity"] = {"requiredDuringSchedulingIgnoredDuringExecution": [
                {"labelSelector": {"matchLabels": {"app": name}}, "topologyKey": "kubernetes.io/hostname"}]}
        if rng.chance(10):
            other = names[rng.below(n_workloads)]
            aff.setdefault("podAffinity", {})["preferredDuringSchedulingIgnoredDuringExecution"] = [
                {"weight": 1 + rng.below(100),
                 "podAffinityTerm": {"labelSelector": {"matchLabels": {"app": other}},
                                     "namespaces": [f"ns-{names.index(other) % 8}"],
                                     "topologyKey": "topology.kubernetes.io/zone"}}]
        if aff:
            extra["affinity"] = aff
        if rng.chance(10):
            extra["topologySpreadConstraints"] = [
                {"maxSkew": 1 + rng.below(5), "topologyKey": "topology.kubernetes.io/zone",
                 "whenUnsatisfiable": rng.pick(["DoNotSchedule", "ScheduleAnyway"]),
                 "labelSelector": {"matchLabels": {"app": name}}}]
        apps[w % n_apps].Resource.Deployments.append(_deployment(name, ns, replicas, labels, cpu, mem, extra))
        cluster.Services.append(_service(name, ns, {"app": name}))
    return cluster, apps


def make_c4(n_nodes: int = 2000, n_workloads: int = 50, replicas: int = 100, fill_pct: int = 85, seed_no: int = 4):
    """C4: base cluster ~fill_pct % full + pending pods that do not all fit; 8 candidate node specs.
    Returns (cluster, apps, specs)."""
    rng = SplitMix64(SEED_BASE + seed_no)
    cluster = ResourceTypes()
    for i in range(n_nodes):
        cores = rng.pick(CPU_SHAPES)
        eph = rng.pick([100, 200, 500, 1000, 2048])
        cluster.Nodes.append(_node(f"node-{i:05d}", cores, f"zone-{i % 16:02d}", "", eph))
        # one running "filler" pod per node occupying ~fill_pct % of cpu and memory
        cpu_m = cores * 10 * fill_pct
        mem_mi = cores * 4 * 1024 * fill_pct // 100
        cluster.Pods.append({"apiVersion": "v1", "kind": "Pod",
                             "metadata": {"name": f"filler-{i:05d}", "namespace": "kube-system", "labels": {"app": "filler"}},
                             "spec": {"nodeName": f"node-{i:05d}",
                                      "containers": [{"name": "c", "image": "registry.local/filler:v1",
                                                      "resources": {"requests": {"cpu": f"{cpu_m}m", "memory": f"{mem_mi}Mi"}}}]}})
    app = AppResource("pending", ResourceTypes())
    for w in range(n_workloads):
        name = f"pend-{w:03d}"
        cpu = _log_uniform(rng, 500, 8000, 50)
        mem = _log_uniform(rng, 512, 32768, 64)
        extra = {}
        if rng.chance(20):
            extra["affinity"] = {"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [
                {"labelSelector": {"matchLabels": {"app": name}}, "topologyKey": "kubernetes.io/hostname"}]}}
        app.Resource.Deployments.append(_deployment(name, "default", replicas, {"app": name}, cpu, mem, extra))
        cluster.Services.append(_service(name, "default", {"app": name}))
    specs = []
    for si, cores in enumerate(CPU_SHAPES):
        specs.append(_node(f"spec-{si}", cores, f"zone-{si % 16:02d}", "", 500))
    return cluster, [app], specs
