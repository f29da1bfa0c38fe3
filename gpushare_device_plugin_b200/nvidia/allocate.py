"""Allocate — mirror of pkg/gpu/nvidia/allocate.go. The decision (request decode, candidate filter,
assume-time order, first pod whose limit equals the request, env synthesis, response encode) runs in
the C ABI (gsb_allocate: wire bytes in, wire bytes out); this module does the control-plane I/O
around it: list the node's pending pods, PATCH the matched pod's annotations."""
from __future__ import annotations

import ctypes as C
import logging
import time
from typing import List, Optional, Sequence

from .. import _abi
from .._abi import AllocateCtx, Pod, check, lib
from . import const, podmanager
from .podutils import (getAssumeTimeFromPodAnnotation, getGPUIDFromPodAnnotation, getGPUMemoryFromPodResource,
                       patchPodAnnotationSpecAssigned)

log = logging.getLogger("gpushare.nvidia")

_RESP_CAP = 1 << 16


class AllocateContext:
    """Immutable per-plugin inputs of gsb_allocate (devNameMap, getGPUMemory(), metric, CGPU flag)."""

    def __init__(self, devNameMap: dict, slices: int, unit_gib: bool, disableCGPUIsolation: bool):
        self._uuids_b = [u.encode() for u in devNameMap]
        self._uuids = (C.c_char_p * len(self._uuids_b))(*self._uuids_b)
        self._minors = (C.c_uint32 * len(self._uuids_b))(*devNameMap.values())
        self.ctx = AllocateCtx(self._uuids, self._minors, len(self._uuids_b), slices, 1 if unit_gib else 0,
                               1 if disableCGPUIsolation else 0)


def pod_table(pods: Sequence[dict], nodeName: str):
    """v1.Pod JSON -> gsb_pod[]: the fields the reference reads (podutils.go:37-131, podmanager.go:187-201)."""
    arr = (Pod * len(pods))()
    keep = []  # keep the bytes alive while C reads them
    for i, p in enumerate(pods):
        md = p.get("metadata") or {}
        ann = md.get("annotations") or {}
        name, ns, uid = (md.get("name") or "").encode(), (md.get("namespace") or "").encode(), (md.get("uid") or "").encode()
        keep.append((name, ns, uid))
        e = arr[i]
        e.name, e.ns, e.uid = name, ns, uid
        e.gpu_mem_limit = getGPUMemoryFromPodResource(p)
        e.assume_time = getAssumeTimeFromPodAnnotation(p)
        e.gpu_idx = max(-1, min(getGPUIDFromPodAnnotation(p), 0x7FFFFFFF))
        e.has_assume_time = 1 if const.EnvResourceAssumeTime in ann else 0
        e.has_assigned = 1 if const.EnvAssignedFlag in ann else 0
        e.assigned_is_false = 1 if ann.get(const.EnvAssignedFlag) == "false" else 0
        e.on_node = 1 if (p.get("spec") or {}).get("nodeName") == nodeName else 0
    return arr, keep


def buildErrResponse(actx: AllocateContext, req: bytes) -> bytes:  # allocate.go:24-39
    buf = C.create_string_buffer(_RESP_CAP)
    n = C.c_size_t(0)
    check(lib.gsb_allocate_err_response(C.byref(actx.ctx), req, len(req), buf, _RESP_CAP, C.byref(n)),
          "gsb_allocate_err_response")
    return buf.raw[: n.value]


def decodable(actx: AllocateContext, req: bytes) -> bool:
    """Would gogo's AllocateRequest.Unmarshal (api.pb.go:2141-2221) accept these bytes?"""
    n = C.c_size_t(0)
    return lib.gsb_allocate_err_response(C.byref(actx.ctx), req, len(req), None, 0, C.byref(n)) != _abi.GSB_ERR_MALFORMED


class PendingPodCache:
    """the reference LISTs (and JSON-decodes) every pending pod of the node on
    every Allocate while holding the plugin lock. This keeps the last LIST and its gsb_pod table for
    `ttl` seconds. It is optimistic, never authoritative: a request that finds no candidate in a
    cached table re-LISTs and is decided again on fresh data, a failed PATCH drops the cache, and a pod
    claimed by one request is hidden from the next (its PATCH runs outside the lock) — in every table
    rebuilt from a later LIST too, until the apiserver's own copy stops saying assigned == "false": a LIST
    is a snapshot that can pre-date a PATCH still in flight. ttl == 0 LISTs on every call like the
    reference."""

    def __init__(self, ttl: float):
        self.ttl = ttl
        self.claimed: set = set()  # uids claimed here and not yet confirmed by a LIST
        self.pods: Optional[List[dict]] = None
        self.table = None
        self._keep = None
        self.stamp = 0.0

    def fresh(self) -> bool:
        return self.pods is not None and self.ttl > 0 and (time.monotonic() - self.stamp) < self.ttl

    def load(self, plugin):
        self.pods = podmanager.getPendingPodsInNode(plugin.queryKubelet, plugin.kubeletClient)
        self.table, self._keep = pod_table(self.pods, podmanager.nodeName)
        self.stamp = time.monotonic()
        if self.claimed:  # reconcile: keep in-flight claims hidden, forget confirmed / departed ones
            still = set()
            for i, p in enumerate(self.pods):
                uid = (p.get("metadata") or {}).get("uid") or ""
                if uid in self.claimed and self.table[i].has_assigned and self.table[i].assigned_is_false:
                    self.table[i].assigned_is_false = 0
                    still.add(uid)
            self.claimed = still

    def claim(self, i: int) -> str:
        self.table[i].assigned_is_false = 0  # no longer a candidate (isGPUMemoryAssumedPod)
        uid = (self.pods[i].get("metadata") or {}).get("uid") or ""
        self.claimed.add(uid)
        return uid

    def unclaim(self, uid: str):
        self.claimed.discard(uid)  # the PATCH failed: still unassigned on the apiserver, a candidate again

    def drop(self):
        self.pods = self.table = self._keep = None


def _decide(actx, cache, req: bytes):
    buf = C.create_string_buffer(_RESP_CAP)
    n, pod_index, pod_req = C.c_size_t(0), C.c_int32(-1), C.c_uint32(0)
    kind = lib.gsb_allocate(C.byref(actx.ctx), cache.table, len(cache.pods), req, len(req), buf, _RESP_CAP,
                            C.byref(n), C.byref(pod_index), C.byref(pod_req))
    return kind, buf.raw[: n.value] if kind > 0 else b"", pod_index.value, pod_req.value


def allocate(plugin, req: bytes) -> bytes:
    """NvidiaDevicePlugin.Allocate (allocate.go:42-198). Never raises: every failure is encoded in the
    response envs, exactly like the reference (gRPC status stays OK)."""
    log.info("----Allocating GPU for gpu mem is started----")
    actx: AllocateContext = plugin.allocate_ctx
    cache: PendingPodCache = plugin.pod_cache
    with plugin.lock:  # allocate.go:59-60 — held for list + decide + claim only, not for the PATCH
        log.info("checking...")
        try:
            was_cached = cache.fresh()
            if not was_cached:
                cache.load(plugin)
            kind, resp, pidx, pod_req = _decide(actx, cache, req)
            if kind > 0 and kind != _abi.GSB_ALLOC_MATCHED and was_cached:
                # the cache may simply be older than the pod being started: the error response AND, on a one-GPU node,
                # the single-GPU shortcut (allocate.go:151-177) are re-decided on a fresh LIST, as the reference —
                # which LISTs on every call — would have decided them
                cache.load(plugin)
                kind, resp, pidx, pod_req = _decide(actx, cache, req)
        except Exception as e:  # noqa: BLE001  allocate.go:62-66
            cache.drop()
            log.info("invalid allocation requst: Failed to find candidate pods due to %s", e)
            return buildErrResponse(actx, req)
        if kind < 0:
            log.warning("Allocate: %s", _abi.last_error() or lib.gsb_strerror(kind).decode())
            return b""  # undecodable request: nothing to answer for
        log.info("RequestPodGPUs: %d", pod_req)
        pod = None
        if kind == _abi.GSB_ALLOC_MATCHED:
            pod = cache.pods[pidx]
            claimed_uid = cache.claim(pidx)
    if kind == _abi.GSB_ALLOC_MATCHED:
        md = pod["metadata"]
        log.info("Found Assumed GPU shared Pod %s in ns %s with GPU Memory %d", md.get("name"), md.get("namespace"),
                 pod_req)
        body = patchPodAnnotationSpecAssigned()  # allocate.go:131
        err = None
        try:
            podmanager.clientset.patch_pod(md.get("namespace"), md.get("name"), body)
        except Exception as e:  # noqa: BLE001
            err = e
            if str(e) == const.OptimisticLockErrorMsg:  # allocate.go:138-144: one retry
                try:
                    podmanager.clientset.patch_pod(md.get("namespace"), md.get("name"), body)
                    err = None
                except Exception as e2:  # noqa: BLE001
                    err = e2
        if err is not None:
            log.warning("Failed due to %s", err)
            with plugin.lock:
                cache.unclaim(claimed_uid)
                cache.drop()
            return buildErrResponse(actx, req)
        log.info("----Allocating GPU for gpu mem for %s is ended----", md.get("name"))
    elif kind == _abi.GSB_ALLOC_SINGLE_GPU:
        log.info("this node has only one gpu device,skip to search pod and directly specify the device")
    else:
        log.warning("invalid allocation requst: request GPU memory %d can't be satisfied.", pod_req)
    return resp
