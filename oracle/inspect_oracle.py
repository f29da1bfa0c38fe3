"""inspect_oracle.py — CPU restatement of cmd/inspect (kubectl-inspect-gpushare) for parity tests. TEST
INFRASTRUCTURE: only tests/ may import it; the product (gpushare_device_plugin_b200/cmd/inspect.py) never does.

An independent second reading of the reference, written without looking at the product's structure:
    cmd/inspect/main.go:31-74        which nodes / pods are fetched (done by the caller here: plain lists in)
    cmd/inspect/podinfo.go:95-134    filterActivePods, gpuMemoryInPod
    cmd/inspect/nodeinfo.go:46-271   buildAllNodeInfos, buildNodeInfoWithPods, buildDeviceInfo, getDeivceInfo, setUnit,
                                     GetAllocation (json.Unmarshal into map[int]map[string]int, then strconv.Atoi of ids)
    cmd/inspect/display.go:15-245    displayDetails, displaySummary
    text/tabwriter (Go standard library, go1.10; NOT under /root/reference): NewWriter(out, 0, 0, 2, ' ', 0) — restated
                                     from its published algorithm in `tabwrite` below, non-recursively
PARITY UNPINNED: no output of the reference exists (no Go toolchain here, no fixtures in the reference).

Go-map iteration makes two orders of the reference's output random: the order of nodes (nodeinfo.go:130-133 ranges over
a map) and, in -d, the order in which a node's devices — hence its pod rows — are visited (display.go:60). This oracle,
like the product, fixes both to ascending order (node name; device index, the pending pseudo-device -1 first); the
reference's output is always a permutation of those rows.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional

from oracle.wire_oracle import quantity_value

RESOURCE, COUNT = "aliyun.com/gpu-mem", "aliyun.com/gpu-count"   # main.go:11-12
IDX_ANN = "ALIYUN_COM_GPU_MEM_IDX"                                # main.go:19
ALLOC_ANN = "scheduler.framework.gpushare.allocation"             # main.go:22


# ------------------------------------------------------------------ text/tabwriter, minwidth 0, tabwidth 0, padding 2, ' ', flags 0
def tabwrite(text: str, padding: int = 2) -> str:
    """Cells are tab-TERMINATED; the text after a line's last tab is a trailing cell that belongs to no column. A
    column block is a maximal run of consecutive lines that all have a cell in that column; its width is the widest
    cell of the run plus the padding; cells are left-aligned and padded with spaces. (Go states this recursively —
    format() opens a block for column k inside a block of column k-1 — which is the same set of runs, since a line
    with a cell in column k has one in column k-1.)"""
    assert text.endswith("\n") or text == ""
    rows = [ln.split("\t") for ln in text.split("\n")[:-1]]
    n_term = [len(r) - 1 for r in rows]
    out = []
    for i, r in enumerate(rows):
        line = ""
        for k in range(n_term[i]):
            lo = i
            while lo > 0 and n_term[lo - 1] > k:
                lo -= 1
            hi = i
            while hi + 1 < len(rows) and n_term[hi + 1] > k:
                hi += 1
            width = max(len(rows[j][k]) for j in range(lo, hi + 1)) + padding
            line += r[k] + " " * (width - len(r[k]))
        out.append(line + r[-1])
    return "".join(ln + "\n" for ln in out)


# ------------------------------------------------------------------ Go parsing rules
def go_atoi(s) -> Optional[int]:  # strconv.Atoi: optional sign, decimal digits, int64 range
    if not isinstance(s, str):
        return None
    body = s[1:] if s[:1] in "+-" else s
    if not body or not all("0" <= c <= "9" for c in body):
        return None
    v = int(s)
    return v if -(1 << 63) <= v < (1 << 63) else None


class _GoJsonError(Exception):
    pass


def _go_int_literal(tok: str) -> int:
    """encoding/json into a Go int: the number's LITERAL must be a base-10 integer (no fraction, no exponent)."""
    v = go_atoi(tok)
    if v is None or tok.startswith("+"):
        raise _GoJsonError(tok)
    return v


def get_allocation(pod: dict) -> Dict[int, int]:  # nodeinfo.go:244-271
    ann = (pod.get("metadata") or {}).get("annotations")
    if ann is None or ALLOC_ANN not in ann:
        return {}
    try:  # json.Unmarshal([]byte(s), &map[int]map[string]int{})
        doc = json.loads(ann[ALLOC_ANN], parse_int=lambda t: ("int", t), parse_float=lambda t: ("float", t))
        if doc is None:
            return {}
        if not isinstance(doc, dict):
            raise _GoJsonError("not an object")
        allocation: Dict[int, Dict[str, int]] = {}
        for key, inner in doc.items():
            k = go_atoi(key)  # map[int] keys: strconv.ParseInt(key, 10, 64)
            if k is None:
                raise _GoJsonError(key)
            if inner is None:
                allocation[k] = {}
                continue
            if not isinstance(inner, dict):
                raise _GoJsonError("inner not an object")
            m = {}
            for id_, val in inner.items():
                if val is None:
                    m[id_] = 0
                elif isinstance(val, tuple) and val[0] == "int":
                    m[id_] = _go_int_literal(val[1])
                else:
                    raise _GoJsonError("value is not an integer literal")
            allocation[k] = m
    except (ValueError, _GoJsonError):
        return {}
    out: Dict[int, int] = {}
    for inner in allocation.values():
        for id_, mem in inner.items():
            idx = go_atoi(id_)
            if idx is None:
                return {}  # nodeinfo.go:263-266
            out[idx] = out.get(idx, 0) + mem
    return out


def gpu_memory_in_pod(pod: dict) -> int:  # podinfo.go:124-134, display.go:247-255
    total = 0
    for c in (pod.get("spec") or {}).get("containers") or []:
        limits = (c.get("resources") or {}).get("limits") or {}
        if RESOURCE in limits:
            total += quantity_value(limits[RESOURCE])
    return total


def filter_active_pods(pods: List[dict]) -> List[dict]:  # podinfo.go:95-106
    return [p for p in pods if (p.get("status") or {}).get("phase") not in ("Succeeded", "Failed")]


def is_gpu_sharing_node(node: dict) -> bool:  # nodeinfo.go:213-221
    alloc = (node.get("status") or {}).get("allocatable") or {}
    return RESOURCE in alloc and quantity_value(alloc[RESOURCE]) > 0


# ------------------------------------------------------------------ nodeinfo.go
class _Node:
    def __init__(self, node: dict):
        alloc = (node.get("status") or {}).get("allocatable") or {}
        self.node = node
        self.name = (node.get("metadata") or {}).get("name") or ""
        self.pods: List[dict] = []
        self.gpu_count = quantity_value(alloc[COUNT]) if COUNT in alloc else 0       # nodeinfo.go:85-93
        self.total = quantity_value(alloc[RESOURCE]) if RESOURCE in alloc else 0      # nodeinfo.go:75-83
        # devs: idx -> [used, total, pods]   (nodeinfo.go:111-119)
        self.devs: Dict[int, list] = {i: [0, self.total // self.gpu_count, []] for i in range(self.gpu_count)}

    def address(self) -> str:  # display.go:24-33, 170-179
        for a in (self.node.get("status") or {}).get("addresses") or []:
            if a.get("type") == "InternalIP":
                return a.get("address") or ""
        return "unknown"

    def device_info(self, pod: dict) -> Dict[int, int]:  # getDeivceInfo, nodeinfo.go:168-196
        allocation = get_allocation(pod)
        if len(allocation) != 0:
            return allocation
        id_ = -1
        ann = (pod.get("metadata") or {}).get("annotations") or {}
        if len(ann) > 0 and IDX_ANN in ann:
            v = go_atoi(ann[IDX_ANN])
            id_ = v if v is not None else -1
        return {id_: gpu_memory_in_pod(pod)}

    def build_device_info(self) -> None:  # nodeinfo.go:142-166
        per_dev = self.total // self.gpu_count if self.gpu_count > 0 else 0
        for pod in self.pods:
            if gpu_memory_in_pod(pod) <= 0:
                continue
            for dev_id, used in self.device_info(pod).items():
                if dev_id not in self.devs:
                    self.devs[dev_id] = [0, per_dev, []]
                self.devs[dev_id][0] += used
                self.devs[dev_id][2].append(pod)

    def dev_string(self, idx: int) -> str:  # DeviceInfo.String, nodeinfo.go:22-27
        used, total, _ = self.devs[idx]
        return "%d" % used if idx == -1 else "%d/%d" % (used, total)


def build_all_node_infos(pods: List[dict], nodes: List[dict]):  # nodeinfo.go:46-134
    by_name: Dict[str, _Node] = {}
    for node in nodes:
        name = (node.get("metadata") or {}).get("name") or ""
        info = by_name.get(name)
        if info is None:
            info = by_name[name] = _Node(node)
        for pod in pods:  # (a node listed twice gets its pods twice: nodeinfo.go:123-127 sits outside the else)
            if (pod.get("spec") or {}).get("nodeName") == name:
                info.pods.append(pod)
    infos = [by_name[k] for k in sorted(by_name)]  # Go ranges over the map: any order; ascending here
    unit = ""
    for info in infos:
        if info.total > 0:
            if unit == "" and info.gpu_count != 0:  # setUnit, nodeinfo.go:227-243: first node decides, for the process
                unit = "MiB" if info.total // info.gpu_count > 100 else "GiB"
            info.build_device_info()
    return infos, unit


# ------------------------------------------------------------------ display.go
def display_summary(infos, unit: str) -> str:  # display.go:141-245
    has_pending = any(-1 in n.devs for n in infos)
    max_gpu = max([n.gpu_count for n in infos] + [0])
    text = "NAME\tIPADDRESS\t" + "".join("GPU%d(Allocated/Total)\t" % i for i in range(max_gpu))
    if has_pending:
        text += "PENDING(Allocated)\t"
    text += "GPU Memory(%s)\n" % unit
    used_cluster = total_cluster = line_len = 0
    for n in infos:
        if n.total <= 0:
            continue
        used, cells = 0, []
        for i in range(max_gpu):
            if i in n.devs:
                cells.append(n.dev_string(i))
                used += n.devs[i][0]
            else:
                cells.append("0/0")
        pending = ""
        if -1 in n.devs:
            pending = "%d" % n.devs[-1][0]
            used += n.devs[-1][0]
        row = "%s\t%s\t" % (n.name, n.address()) + "".join(c + "\t" for c in cells)
        if has_pending:
            row += pending + "\t"
        row += "%d/%d\n" % (used, n.total)
        text += row
        if line_len == 0:
            line_len = len(row.encode()) + 20  # buf.Len() + 20, display.go:216-218
        used_cluster += used
        total_cluster += n.total
    text += "-" * line_len + "\n"
    text += "Allocated/Total GPU Memory In Cluster:\n"
    usage = used_cluster / total_cluster * 100 if total_cluster > 0 else 0.0
    text += "%d/%d (%d%%)\t\n" % (used_cluster, total_cluster, int(usage))
    return tabwrite(text)


def display_details(infos) -> str:  # display.go:15-129
    text = ""
    used_cluster = total_cluster = line_len = 0
    for n in infos:
        if n.total <= 0:
            continue
        text += "\nNAME:\t%s\nIPADDRESS:\t%s\n\n" % (n.name, n.address())
        text += "NAME\tNAMESPACE\t" + "".join("GPU%d(Allocated)\t" % i for i in range(n.gpu_count))
        pending = -1 in n.devs
        if pending:
            text += "Pending(Allocated)\t"
        text += "\n"
        used, rows, seen = 0, "", set()
        for i in sorted(n.devs):  # Go ranges over the map: any order; ascending here
            dev_used, _, dev_pods = n.devs[i]
            used += dev_used
            for pod in dev_pods:
                md = pod.get("metadata") or {}
                if md.get("uid") in seen:
                    continue
                rows += "%s\t%s\t" % (md.get("name") or "", md.get("namespace") or "")
                for k in range(n.gpu_count + (1 if pending else 0)):
                    allocation = get_allocation(pod)
                    if len(allocation) != 0:
                        rows += "%d\t" % allocation.get(k, 0)
                    elif k == i or (i == -1 and k == n.gpu_count):
                        rows += "%d\t" % gpu_memory_in_pod(pod)
                    else:
                        rows += "0\t"
                rows += "\n"
                seen.add(md.get("uid"))
        if line_len == 0:
            line_len = len(rows.encode()) + 10
        text += rows
        text += "Allocated :\t%d (%d%%)\t\n" % (used, int(used / n.total * 100))
        text += "Total :\t%d \t\n" % n.total
        text += "-" * line_len + "\n"
        total_cluster += n.total
        used_cluster += used
    text += "\n\n"
    usage = used_cluster / total_cluster * 100 if total_cluster > 0 else 0.0
    text += "Allocated/Total GPU Memory In Cluster:\t%d/%d (%d%%)\t\n" % (used_cluster, total_cluster, int(usage))
    return tabwrite(text)


def inspect(all_nodes: List[dict], all_pods: List[dict], node_name: str = "", details: bool = False) -> str:
    """main.go:31-74 over plain lists: all_nodes / all_pods are what the apiserver holds."""
    if node_name == "":
        nodes = [n for n in all_nodes if is_gpu_sharing_node(n)]
        pods = filter_active_pods(all_pods)
    else:
        nodes = [n for n in all_nodes if (n.get("metadata") or {}).get("name") == node_name]
        pods = filter_active_pods([p for p in all_pods if (p.get("spec") or {}).get("nodeName") == node_name])
    infos, unit = build_all_node_infos(pods, nodes)
    return display_details(infos) if details else display_summary(infos, unit)
