"""Node/pod control-plane helpers — mirror of pkg/gpu/nvidia/podmanager.go."""
from __future__ import annotations

import logging
import os
import time
from typing import List, Optional

from . import const, kubeclient
from .podutils import getAssumeTimeFromPodAnnotation, isGPUMemoryAssumedPod

log = logging.getLogger("gpushare.nvidia")

clientset: Optional[kubeclient.Clientset] = None
nodeName: str = ""
retries = 8  # podmanager.go:26


def kubeInit(cs: Optional[kubeclient.Clientset] = None, node: Optional[str] = None) -> None:
    """podmanager.go:29-57. Tests inject a clientset; production reads $KUBECONFIG / in-cluster."""
    global clientset, nodeName
    clientset = cs if cs is not None else kubeclient.from_environment()
    nodeName = node if node is not None else os.environ.get("NODE_NAME", "")
    if nodeName == "":
        log.critical("Please set env NODE_NAME")
        raise SystemExit(1)


def disableCGPUIsolationOrNot() -> bool:  # podmanager.go:59-72
    node = clientset.get_node(nodeName)
    labels = (node.get("metadata") or {}).get("labels") or {}
    if labels.get(const.EnvNodeLabelForDisableCGPU) == "true":
        log.info("enable gpusharing mode and disable cgpu mode")
        return True
    return False


def patchGPUCount(gpuCount: int) -> None:
    """podmanager.go:74-99 + nodeutil.PatchNodeStatus (vendor/k8s.io/kubernetes/pkg/util/node/node.go:149-182):
    a two-way strategic merge patch of the node status that sets capacity/allocatable
    aliyun.com/gpu-count. resource.NewQuantity(n, DecimalSI) marshals as the string "<n>"."""
    node = clientset.get_node(nodeName)
    cap = (node.get("status") or {}).get("capacity") or {}
    if const.resourceCount in cap and str(cap[const.resourceCount]) == str(gpuCount):
        log.info("No need to update Capacity %s", const.resourceCount)
        return
    patch = {"status": {"allocatable": {const.resourceCount: str(gpuCount)},
                        "capacity": {const.resourceCount: str(gpuCount)}}}
    try:
        clientset.patch_node_status(nodeName, patch)
        log.info("Updated Capacity %s successfully.", const.resourceCount)
    except Exception:
        log.info("Failed to update Capacity %s.", const.resourceCount)
        raise


def getPodList(kubeletClient) -> dict:  # podmanager.go:101-123
    podList = kubeletClient.GetNodeRunningPods()
    items = [p for p in podList.get("items") or [] if (p.get("status") or {}).get("phase") == "Pending"]
    if len(items) == 0:
        raise LookupError("not found pending pod")
    return {"items": items}


def getPodListsByListAPIServer() -> dict:  # podmanager.go:142-160
    selector = f"spec.nodeName={nodeName},status.phase=Pending"
    err = None
    for attempt in range(4):  # one try + up to 3 retries, 1 s apart
        try:
            return clientset.list_pods(selector)
        except Exception as e:  # noqa: BLE001
            err = e
            if attempt < 3:
                time.sleep(1)
    raise RuntimeError(f"failed to get Pods assigned to node {nodeName}") from err


def getPodListsByQueryKubelet(kubeletClient) -> dict:  # podmanager.go:125-140
    err = None
    for attempt in range(retries + 1):
        try:
            return getPodList(kubeletClient)
        except Exception as e:  # noqa: BLE001
            err = e
            if attempt < retries:
                log.warning("failed to get pending pod list, retry")
                time.sleep(0.1)
    log.warning("not found from kubelet /pods api, start to list apiserver (%s)", err)
    return getPodListsByListAPIServer()


def getPendingPodsInNode(queryKubelet: bool, kubeletClient) -> List[dict]:  # podmanager.go:162-212
    podList = getPodListsByQueryKubelet(kubeletClient) if queryKubelet else getPodListsByListAPIServer()
    pods, seen = [], set()
    for pod in podList.get("items") or []:
        if (pod.get("spec") or {}).get("nodeName") != nodeName:
            log.warning("Pod name %s in ns %s is not assigned to node %s as expected", pod["metadata"].get("name"),
                        pod["metadata"].get("namespace"), nodeName)
            continue
        uid = pod["metadata"].get("uid")
        if uid not in seen:
            pods.append(pod)
            seen.add(uid)
    return pods


def getCandidatePods(queryKubelet: bool, client) -> List[dict]:
    """podmanager.go:215-262. The product's Allocate does filter+sort+match inside gsb_allocate (C ABI);
    this Python form exists for callers that want the list itself (same order rule)."""
    cand = [p for p in getPendingPodsInNode(queryKubelet, client) if isGPUMemoryAssumedPod(p)]
    return makePodOrderdByAge(cand)


def makePodOrderdByAge(pods: List[dict]) -> List[dict]:
    """sort.Sort with Less = `<=` (podmanager.go:241-262): Go 1.10's small-slice path restated
    (ShellSort gap 6 for <= 12 elements, then insertion sort) — see csrc/gsb_wire.cc."""
    order = list(pods)
    key = getAssumeTimeFromPodAnnotation
    n = len(order)
    if 1 < n <= 12:
        for i in range(6, n):
            if key(order[i]) <= key(order[i - 6]):
                order[i], order[i - 6] = order[i - 6], order[i]
    for i in range(1, n):
        j = i
        while j > 0 and key(order[j]) <= key(order[j - 1]):
            order[j], order[j - 1] = order[j - 1], order[j]
            j -= 1
    return order
