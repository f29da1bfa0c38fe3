"""kubectl-inspect-gpushare — mirror of cmd/inspect/{main,podinfo,nodeinfo,display}.go.

    python -m gpushare_device_plugin_b200.cmd.inspect [-d] [node]

Host-only: lists nodes whose allocatable aliyun.com/gpu-mem > 0 and the active pods from the
kube-APISERVER (via $KUBECONFIG, podinfo.go:27-46 — not from the kubelet) and prints the per-GPU
allocated/total table through a text/tabwriter restatement. Quirks are reproduced, not fixed: the unit
banner says "MiB" as soon as a GPU has more than 100 units (nodeinfo.go:227-243), so a 179-GiB-slice
B200 is labelled MiB; Go map iteration order (node order, device order in -d) is replaced by a sorted
order, the one deterministic choice the reference's output is a permutation of."""
from __future__ import annotations

import json
import os
import sys
import time
from typing import Dict, List, Optional

from ..nvidia import kubeclient
from ..nvidia.podutils import quantityValue
from .tabwriter import Writer

resourceName = "aliyun.com/gpu-mem"
countName = "aliyun.com/gpu-count"
envNVGPUID = "ALIYUN_COM_GPU_MEM_IDX"
gpushareAllocationFlag = "scheduler.framework.gpushare.allocation"
retries = 5

memoryUnit = ""


def _atoi(s: str) -> Optional[int]:
    body = s[1:] if s[:1] in "+-" else s
    return int(s) if body and all("0" <= c <= "9" for c in body) else None


def gpuMemoryInPod(pod: dict) -> int:  # podinfo.go:124-134 == display.go:247-255
    total = 0
    for c in (pod.get("spec") or {}).get("containers") or []:
        limits = (c.get("resources") or {}).get("limits") or {}
        if resourceName in limits:
            total += quantityValue(limits[resourceName])
    return total


def _parse_int64(s) -> Optional[int]:  # strconv.ParseInt(s, 10, 64) / strconv.Atoi
    v = _atoi(s) if isinstance(s, str) else None
    return v if v is not None and -(1 << 63) <= v < (1 << 63) else None


def GetAllocation(pod: dict) -> Dict[int, int]:  # nodeinfo.go:244-271
    """json.Unmarshal into map[int]map[string]int, then strconv.Atoi of every inner key. ANY decoding error — an outer
    key that is not an integer, a value whose JSON literal is not a plain integer (2.0 and 2e0 are refused for a Go
    int), a non-object where an object is expected — makes the reference return the empty map, and the caller then
    falls back to the IDX annotation (nodeinfo.go:168-196). null is accepted where Go accepts it (nil map / zero)."""
    ann = (pod.get("metadata") or {}).get("annotations")
    if ann is None or gpushareAllocationFlag not in ann:
        return {}
    try:
        doc = json.loads(ann[gpushareAllocationFlag], parse_int=lambda lit: ("i", lit), parse_float=lambda lit: ("f", lit))
    except ValueError:
        return {}
    if doc is None:
        return {}
    if not isinstance(doc, dict):
        return {}
    decoded = []
    for outer_key, containerAllocation in doc.items():
        if _parse_int64(outer_key) is None:
            return {}
        if containerAllocation is None:
            continue
        if not isinstance(containerAllocation, dict):
            return {}
        for id_, gpuMem in containerAllocation.items():
            if gpuMem is None:
                decoded.append((id_, 0))
            elif isinstance(gpuMem, tuple) and gpuMem[0] == "i" and not gpuMem[1].startswith("+") and _parse_int64(gpuMem[1]) is not None:
                decoded.append((id_, int(gpuMem[1])))
            else:
                return {}
    out: Dict[int, int] = {}
    for id_, gpuMem in decoded:
        idx = _parse_int64(id_)
        if idx is None:
            return {}  # nodeinfo.go:263-266
        out[idx] = out.get(idx, 0) + gpuMem
    return out


class DeviceInfo:
    def __init__(self, idx: int, totalGPUMem: int):
        self.idx, self.pods, self.usedGPUMem, self.totalGPUMem = idx, [], 0, totalGPUMem

    def __str__(self):  # nodeinfo.go:22-27
        return f"{self.usedGPUMem}" if self.idx == -1 else f"{self.usedGPUMem}/{self.totalGPUMem}"


class NodeInfo:
    def __init__(self, node: dict):
        alloc = (node.get("status") or {}).get("allocatable") or {}
        self.node, self.pods = node, []
        self.gpuCount = quantityValue(alloc[countName]) if countName in alloc else 0
        self.gpuTotalMemory = quantityValue(alloc[resourceName]) if resourceName in alloc else 0
        self.devs: Dict[int, DeviceInfo] = {}
        for i in range(self.gpuCount):  # nodeinfo.go:111-119
            self.devs[i] = DeviceInfo(i, self.gpuTotalMemory // self.gpuCount)

    def hasPendingGPUMemory(self) -> bool:
        return -1 in self.devs

    def getDeivceInfo(self, pod: dict) -> Dict[int, int]:  # nodeinfo.go:168-196
        allocation = GetAllocation(pod)
        if allocation:
            return allocation
        id_ = -1
        ann = (pod.get("metadata") or {}).get("annotations") or {}
        if ann and envNVGPUID in ann:
            v = _atoi(ann[envNVGPUID])
            id_ = v if v is not None else -1
        return {id_: gpuMemoryInPod(pod)}

    def buildDeviceInfo(self) -> None:  # nodeinfo.go:142-166
        total = self.gpuTotalMemory // self.gpuCount if self.gpuCount > 0 else 0
        for pod in self.pods:
            if gpuMemoryInPod(pod) <= 0:
                continue
            for devID, used in self.getDeivceInfo(pod).items():
                if devID not in self.devs:
                    self.devs[devID] = DeviceInfo(devID, total)
                self.devs[devID].usedGPUMem += used
                self.devs[devID].pods.append(pod)

    def address(self) -> str:
        for a in (self.node.get("status") or {}).get("addresses") or []:
            if a.get("type") == "InternalIP":
                return a.get("address")
        return "unknown"


def setUnit(gpuMemory: int, gpuCount: int) -> None:  # nodeinfo.go:227-243
    global memoryUnit
    if memoryUnit != "" or gpuCount == 0:
        return
    memoryUnit = "MiB" if gpuMemory // gpuCount > 100 else "GiB"


def buildAllNodeInfos(allPods: List[dict], nodes: List[dict]) -> List[NodeInfo]:  # nodeinfo.go:46-134
    infos: Dict[str, NodeInfo] = {}
    for node in nodes:
        name = node["metadata"]["name"]
        info = infos.get(name)
        if info is None:
            info = infos[name] = NodeInfo(node)
        for pod in allPods:
            if (pod.get("spec") or {}).get("nodeName") == name:
                info.pods.append(pod)
    out = [infos[k] for k in sorted(infos)]  # the reference ranges over a Go map here
    for info in out:
        if info.gpuTotalMemory > 0:
            setUnit(info.gpuTotalMemory, info.gpuCount)
            info.buildDeviceInfo()
    return out


def filterActivePods(pods: List[dict]) -> List[dict]:  # podinfo.go:95-106
    return [p for p in pods if (p.get("status") or {}).get("phase") not in ("Succeeded", "Failed")]


def displaySummary(nodeInfos: List[NodeInfo]) -> str:  # display.go:141-245
    w = Writer(0, 0, 2, " ", 0)
    hasPending = any(n.hasPendingGPUMemory() for n in nodeInfos)
    maxGPU = max([n.gpuCount for n in nodeInfos] + [0])
    head = "NAME\tIPADDRESS\t" + "".join(f"GPU{i}(Allocated/Total)\t" for i in range(maxGPU))
    if hasPending:
        head += "PENDING(Allocated)\t"
    w.write(head + f"GPU Memory({memoryUnit})\n")
    used_c = total_c = prtLineLen = 0
    for n in nodeInfos:
        if n.gpuTotalMemory <= 0:
            continue
        used, cells = 0, []
        for i in range(maxGPU):
            if i in n.devs:
                cells.append(str(n.devs[i]))
                used += n.devs[i].usedGPUMem
            else:
                cells.append("0/0")
        pending = ""
        if -1 in n.devs:
            pending = f"{n.devs[-1].usedGPUMem}"
            used += n.devs[-1].usedGPUMem
        buf = f"{n.node['metadata']['name']}\t{n.address()}\t" + "".join(c + "\t" for c in cells)
        if hasPending:
            buf += pending + "\t"
        buf += f"{used}/{n.gpuTotalMemory}\n"
        w.write(buf)
        if prtLineLen == 0:
            prtLineLen = len(buf.encode()) + 20
        used_c += used
        total_c += n.gpuTotalMemory
    w.write("-" * prtLineLen + "\n")
    w.write("Allocated/Total GPU Memory In Cluster:\n")
    usage = used_c / total_c * 100 if total_c > 0 else 0
    w.write(f"{used_c}/{total_c} ({int(usage)}%)\t\n")
    return w.flush()


def displayDetails(nodeInfos: List[NodeInfo]) -> str:  # display.go:15-129
    w = Writer(0, 0, 2, " ", 0)
    used_c = total_c = prtLineLen = 0
    for n in nodeInfos:
        if n.gpuTotalMemory <= 0:
            continue
        w.write("\n")
        w.write(f"NAME:\t{n.node['metadata']['name']}\n")
        w.write(f"IPADDRESS:\t{n.address()}\n")
        w.write("\n")
        head = "NAME\tNAMESPACE\t" + "".join(f"GPU{i}(Allocated)\t" for i in range(n.gpuCount))
        if n.hasPendingGPUMemory():
            head += "Pending(Allocated)\t"
        w.write(head + "\n")
        used, rows, seen = 0, "", set()
        for i in sorted(n.devs):  # Go map order in the reference
            dev = n.devs[i]
            used += dev.usedGPUMem
            for pod in dev.pods:
                uid = pod["metadata"].get("uid")
                if uid in seen:
                    continue
                rows += f"{pod['metadata'].get('name')}\t{pod['metadata'].get('namespace')}\t"
                count = n.gpuCount + (1 if n.hasPendingGPUMemory() else 0)
                for k in range(count):
                    allocation = GetAllocation(pod)
                    if allocation:
                        rows += f"{allocation.get(k, 0)}\t"
                    elif k == i or (i == -1 and k == n.gpuCount):
                        rows += f"{gpuMemoryInPod(pod)}\t"
                    else:
                        rows += "0\t"
                rows += "\n"
                seen.add(uid)
        if prtLineLen == 0:
            prtLineLen = len(rows.encode()) + 10
        w.write(rows)
        w.write(f"Allocated :\t{used} ({int(used / n.gpuTotalMemory * 100)}%)\t\n")
        w.write(f"Total :\t{n.gpuTotalMemory} \t\n")
        w.write("-" * prtLineLen + "\n")
        total_c += n.gpuTotalMemory
        used_c += used
    w.write("\n\n")
    w.write("Allocated/Total GPU Memory In Cluster:\t")
    usage = used_c / total_c * 100 if total_c > 0 else 0
    w.write(f"{used_c}/{total_c} ({int(usage)}%)\t\n")
    return w.flush()


def _list_with_retries(fn):
    err = None
    for attempt in range(retries + 1):  # podinfo.go:57-76, 78-93
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            err = e
            time.sleep(0.1)
    raise err


def run(argv: List[str], clientset=None) -> str:
    global memoryUnit
    memoryUnit = ""
    details = "-d" in argv
    args = [a for a in argv if a != "-d"]
    nodeName = args[0] if args else ""
    cs = clientset or kubeclient.from_environment()
    if nodeName == "":
        nodes = [n for n in cs.request("GET", "/api/v1/nodes").get("items") or []
                 if quantityValue(((n.get("status") or {}).get("allocatable") or {}).get(resourceName, 0)) > 0]
        pods = filterActivePods(_list_with_retries(lambda: cs.request("GET", "/api/v1/pods")).get("items") or [])
    else:
        nodes = [cs.get_node(nodeName)]
        pods = filterActivePods(_list_with_retries(lambda: cs.list_pods(f"spec.nodeName={nodeName}")).get("items") or [])
    infos = buildAllNodeInfos(pods, nodes)
    return displayDetails(infos) if details else displaySummary(infos)


def main() -> None:
    if not os.environ.get("KUBECONFIG") and not os.path.exists(os.path.join(os.environ.get("HOME", ""), ".kube/config")):
        sys.stderr.write("kubeconfig failed to find, please set KUBECONFIG env\n")  # podinfo.go:32-35
        raise SystemExit(255)
    try:
        sys.stdout.write(run(sys.argv[1:]))
    except Exception as e:  # noqa: BLE001
        sys.stdout.write(f"Failed due to {e}")
        raise SystemExit(1)


if __name__ == "__main__":
    main()
