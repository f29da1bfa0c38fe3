"""Pod annotation/resource helpers — mirror of pkg/gpu/nvidia/podutils.go (pods are the JSON dicts
the kube API returns)."""
from __future__ import annotations

import logging
import math
import time
from fractions import Fraction
from typing import Optional

from .. import _abi
from . import const

log = logging.getLogger("gpushare.nvidia")

_SUFFIX = {"Ki": 1 << 10, "Mi": 1 << 20, "Gi": 1 << 30, "Ti": 1 << 40, "Pi": 1 << 50, "Ei": 1 << 60,
           "k": 10 ** 3, "M": 10 ** 6, "G": 10 ** 9, "T": 10 ** 12, "P": 10 ** 15, "E": 10 ** 18}


def quantityValue(q) -> int:
    """resource.Quantity.Value(): integer value, fractions rounded up."""
    if isinstance(q, int):
        return q
    s = str(q).strip()
    for suf, mult in _SUFFIX.items():
        if s.endswith(suf):
            return math.ceil(Fraction(s[: -len(suf)]) * mult)
    if s.endswith("m"):
        return math.ceil(Fraction(s[:-1]) / 1000)
    return math.ceil(Fraction(s))


def patchPodAnnotationSpecAssigned() -> bytes:  # podutils.go:27-35
    import ctypes as C
    buf = C.create_string_buffer(256)
    n = _abi.check(_abi.lib.gsb_patch_assigned_body(time.time_ns(), buf, len(buf)), "gsb_patch_assigned_body")
    return buf.raw[:n]


def _annotations(pod: dict) -> dict:
    return (pod.get("metadata") or {}).get("annotations") or {}


def getGPUIDFromPodAnnotation(pod: dict) -> int:  # podutils.go:37-61 (strconv.Atoi)
    ann = _annotations(pod)
    if len(ann) > 0 and const.EnvResourceIndex in ann:
        value = ann[const.EnvResourceIndex]
        body = value[1:] if value[:1] in ("+", "-") else value
        if body and all("0" <= ch <= "9" for ch in body) and -(1 << 63) <= int(value) < (1 << 63):
            return int(value)
        log.warning("Failed to parse dev id %s for pod %s in ns %s", value, pod["metadata"].get("name"),
                    pod["metadata"].get("namespace"))
    return -1


def getAssumeTimeFromPodAnnotation(pod: dict) -> int:  # podutils.go:64-75 (strconv.ParseUint base 10)
    s: Optional[str] = _annotations(pod).get(const.EnvResourceAssumeTime)
    if s and all("0" <= ch <= "9" for ch in s) and int(s) < (1 << 64):
        return int(s)
    return 0


def getGPUMemoryFromPodResource(pod: dict) -> int:  # podutils.go:122-131 (spec.containers only)
    total = 0
    for c in (pod.get("spec") or {}).get("containers") or []:
        limits = (c.get("resources") or {}).get("limits") or {}
        if const.resourceName in limits:
            total += quantityValue(limits[const.resourceName])
    return total


def isGPUMemoryAssumedPod(pod: dict) -> bool:  # podutils.go:78-119
    if getGPUMemoryFromPodResource(pod) <= 0:
        return False
    ann = _annotations(pod)
    if const.EnvResourceAssumeTime not in ann:
        return False
    return const.EnvAssignedFlag in ann and ann[const.EnvAssignedFlag] == "false"
