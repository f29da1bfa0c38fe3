"""CPU oracle of the host half of the path — TEST INFRASTRUCTURE ONLY.

Line-by-line restatement, in plain Python, of what the reference's Go code computes for
inventory -> fake devices -> wire bytes, the XID filter, and Allocate. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it; the
product never does.

Parity status — PARITY UNPINNED: the reference has NO tests or golden vectors for this path (its only test,
pkg/kubelet/client/client_test.go, asserts nothing — SURVEY.md §4), and its Go sources cannot be
compiled here (no Go toolchain). The known-answer vectors in tests/golden/wire_kat.json are direct
readings of the cited lines (SURVEY.md §8(c) ①-⑨); the byte-level encoders are additionally checked
against google.protobuf's own encoder driven by a descriptor built from the reference's api.proto
field numbers (tests/test_wire.py). Since round 2 the answers are frozen in tests/golden/allocate_cases.json
(oracle/make_wire_golden.py) so that drift shows up in review; DESIGN.md §4a maps every function to its Go lines.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional, Tuple

# ---------------------------------------------------------------- const.go:11-35

resourceName = "aliyun.com/gpu-mem"
resourceCount = "aliyun.com/gpu-count"
serverSockName = "aliyungpushare.sock"
OptimisticLockErrorMsg = ("the object has been modified; please apply your changes to the latest version and "
                          "try again")
envNVGPU = "NVIDIA_VISIBLE_DEVICES"
EnvResourceIndex = "ALIYUN_COM_GPU_MEM_IDX"
EnvResourceByPod = "ALIYUN_COM_GPU_MEM_POD"
EnvResourceByContainer = "ALIYUN_COM_GPU_MEM_CONTAINER"
EnvResourceByDev = "ALIYUN_COM_GPU_MEM_DEV"
EnvAssignedFlag = "ALIYUN_COM_GPU_MEM_ASSIGNED"
EnvResourceAssumeTime = "ALIYUN_COM_GPU_MEM_ASSUME_TIME"
EnvNodeLabelForDisableCGPU = "cgpu.disable.isolation"
GiBPrefix, MiBPrefix = "GiB", "MiB"
Healthy, Unhealthy = "Healthy", "Unhealthy"  # vendor/.../deviceplugin/v1beta1/constants.go:19-22
Version = "v1beta1"                          # constants.go:27

# ---------------------------------------------------------------- nvidia.go:26-45, bindings.go:346-349


def generateFakeDeviceID(realID: str, fakeCounter: int) -> str:  # nvidia.go:26-28
    return "%s-_-%d" % (realID, fakeCounter)


def extractRealDeviceID(fakeDeviceID: str) -> str:  # nvidia.go:30-32
    return fakeDeviceID.split("-_-")[0]


def mib_from_bytes(total_bytes: int) -> int:  # bindings.go:346-349: *totalMem /= 1024 * 1024
    return total_bytes // (1024 * 1024)


def setGPUMemory(raw_mib: int, metric: str) -> int:  # nvidia.go:34-41
    v = raw_mib
    if metric == GiBPrefix:
        v = raw_mib // 1024
    return v


def getDevices(inventory: List[dict], metric: str = GiBPrefix) -> Tuple[List[List[str]], Dict[str, int], int]:
    """nvidia.go:53-89. `inventory` = what nvml.NewDevice returned per index:
    {"uuid", "path": "/dev/nvidia<minor>", "memory_mib"}. Returns ([ID, Health] list, uuid->minor,
    gpuMemory). The process-global gpuMemory is set from the FIRST device only (nvidia.go:70-72)."""
    devs: List[List[str]] = []
    realDevNames: Dict[str, int] = {}
    gpuMemory = 0
    for d in inventory:
        path = d["path"]
        if not path.startswith("/dev/nvidia"):  # fmt.Sscanf(d.Path, "/dev/nvidia%d", &id)  nvidia.go:65
            raise ValueError("input does not match format")
        digits = ""
        for ch in path[len("/dev/nvidia"):]:
            if ch.isdigit():
                digits += ch
            else:
                break
        if not digits:
            raise ValueError("expected integer")
        realDevNames[d["uuid"]] = int(digits)
        if gpuMemory == 0:
            gpuMemory = setGPUMemory(int(d["memory_mib"]), metric)
        for j in range(gpuMemory):
            devs.append([generateFakeDeviceID(d["uuid"], j), Healthy])
    return devs, realDevNames, gpuMemory


# ---------------------------------------------------------------- gogo wire format (api.pb.go)


def sovApi(x: int) -> int:  # api.pb.go sovApi
    n = 1
    while x >= 0x80:
        x >>= 7
        n += 1
    return n


def encodeVarintApi(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _str_field(tag: int, s: str) -> bytes:
    b = s.encode()
    return bytes([tag]) + encodeVarintApi(len(b)) + b


def marshal_Device(ID: str, health: str) -> bytes:  # api.pb.go:824-843
    out = b""
    if len(ID) > 0:
        out += _str_field(0x0A, ID)
    if len(health) > 0:
        out += _str_field(0x12, health)
    return out


def marshal_ListAndWatchResponse(devs: Iterable[Iterable[str]]) -> bytes:  # api.pb.go:794-812
    out = bytearray()
    for ID, health in devs:
        d = marshal_Device(ID, health)
        out += b"\x0a" + encodeVarintApi(len(d)) + d
    return bytes(out)


def marshal_RegisterRequest(version: str, endpoint: str, resource_name: str) -> bytes:  # api.pb.go:730-767
    out = b""
    if version:
        out += _str_field(0x0A, version)
    if endpoint:
        out += _str_field(0x12, endpoint)
    if resource_name:
        out += _str_field(0x1A, resource_name)
    return out


def marshal_AllocateRequest(container_requests: List[List[str]]) -> bytes:  # api.pb.go:900-957
    out = bytearray()
    for ids in container_requests:
        c = b"".join(_str_field(0x0A, s) for s in ids)
        out += b"\x0a" + encodeVarintApi(len(c)) + c
    return bytes(out)


def marshal_ContainerAllocateResponse(envs: Dict[str, str], key_order: Optional[List[str]] = None) -> bytes:
    """api.pb.go:999-1062 with only Envs set. Go iterates the map in random order (`for k := range
    m.Envs`), so the byte string is defined only up to entry order; key_order picks one."""
    out = bytearray()
    for k in (key_order if key_order is not None else sorted(envs)):
        v = envs[k]
        kb, vb = k.encode(), v.encode()
        mapSize = 1 + len(kb) + sovApi(len(kb)) + 1 + len(vb) + sovApi(len(vb))
        out += b"\x0a" + encodeVarintApi(mapSize) + b"\x0a" + encodeVarintApi(len(kb)) + kb
        out += b"\x12" + encodeVarintApi(len(vb)) + vb
    return bytes(out)


def marshal_AllocateResponse(container_envs: List[Dict[str, str]]) -> bytes:  # api.pb.go:969-987
    out = bytearray()
    for envs in container_envs:
        c = marshal_ContainerAllocateResponse(envs)
        out += b"\x0a" + encodeVarintApi(len(c)) + c
    return bytes(out)


# ---- gogo's generated decoder for AllocateRequest, restated (api.pb.go:2141-2221, 2222-2300, skipApi 2991-3089) ------

class UnmarshalError(ValueError):
    """What grpc-go turns into status INTERNAL ("grpc: error unmarshalling request: ...")."""


def _go_varint(b: bytes, i: int) -> Tuple[int, int]:
    """The inlined loop every field read uses: `shift >= 64` => ErrIntOverflowApi, running out => io.ErrUnexpectedEOF."""
    v = shift = 0
    while True:
        if shift >= 64:
            raise UnmarshalError("proto: integer overflow")
        if i >= len(b):
            raise UnmarshalError("unexpected EOF")
        x = b[i]
        i += 1
        v |= (x & 0x7F) << shift
        if x < 0x80:
            return v & 0xFFFFFFFFFFFFFFFF, i
        shift += 7


def _go_int(v: int) -> int:  # int(uint64): two's complement
    return v - (1 << 64) if v >= (1 << 63) else v


def _int32(v: int) -> int:   # int32(wire >> 3)
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def skipApi(b: bytes, depth: int = 0) -> int:  # api.pb.go:2991-3089; returns the number of bytes of one field
    if depth > 64:  # the product refuses deeper group nesting; Go would recurse on (SURVEY: deviation, unreachable in practice)
        raise UnmarshalError("proto: group nesting too deep")
    if not b:
        raise AssertionError("unreachable in the reference too: callers pass at least the key")
    wire, i = _go_varint(b, 0)
    wt = wire & 7
    if wt == 0:
        shift = 0
        while True:
            if shift >= 64:
                raise UnmarshalError("proto: integer overflow")
            if i >= len(b):
                raise UnmarshalError("unexpected EOF")
            i += 1
            if b[i - 1] < 0x80:
                return i
            shift += 7
    if wt == 1:
        return i + 8
    if wt == 2:
        length, i = _go_varint(b, i)
        length = _go_int(length)
        if length < 0:
            raise UnmarshalError("proto: negative length found during unmarshaling")
        return i + length
    if wt == 3:
        while True:
            start = i
            inner, i = _go_varint(b, i)
            if inner & 7 == 4:
                return i
            if start >= len(b):
                raise UnmarshalError("unexpected EOF")
            i = start + skipApi(b[start:], depth + 1)
            if i > len(b):  # Go would slice past the end on the next round: panic -> the RPC fails either way
                raise UnmarshalError("unexpected EOF")
    if wt == 4:
        return i
    if wt == 5:
        return i + 4
    raise UnmarshalError("proto: illegal wireType %d" % wt)


def _unmarshal_message(b: bytes, name: str, field1: str, on_field1) -> None:
    i, l = 0, len(b)
    while i < l:
        pre = i
        wire, i = _go_varint(b, i)
        field, wt = _int32(wire >> 3), wire & 7
        if wt == 4:
            raise UnmarshalError("proto: %s: wiretype end group for non-group" % name)
        if field <= 0:
            raise UnmarshalError("proto: %s: illegal tag %d (wire type %d)" % (name, field, wire))
        if field == 1:
            if wt != 2:
                raise UnmarshalError("proto: wrong wireType = %d for field %s" % (wt, field1))
            n, i = _go_varint(b, i)
            n = _go_int(n)
            if n < 0:
                raise UnmarshalError("proto: negative length found during unmarshaling")
            if i + n > l:
                raise UnmarshalError("unexpected EOF")
            on_field1(b[i:i + n])
            i += n
        else:
            skippy = skipApi(b[pre:])
            if pre + skippy > l:
                raise UnmarshalError("unexpected EOF")
            i = pre + skippy


def unmarshal_AllocateRequest(b: bytes) -> List[List[bytes]]:
    """AllocateRequest.Unmarshal + ContainerAllocateRequest.Unmarshal: devicesIDs per container, or UnmarshalError."""
    out: List[List[bytes]] = []

    def container(payload: bytes):
        ids: List[bytes] = []
        _unmarshal_message(payload, "ContainerAllocateRequest", "DevicesIDs", ids.append)
        out.append(ids)
    _unmarshal_message(b, "AllocateRequest", "ContainerRequests", container)
    return out


def _read_varint(b: bytes, i: int) -> Tuple[int, int]:
    v = shift = 0
    while True:
        x = b[i]
        i += 1
        v |= (x & 0x7F) << shift
        if not x & 0x80:
            return v, i
        shift += 7


def _fields(b: bytes) -> List[Tuple[int, int, object]]:
    out, i = [], 0
    while i < len(b):
        key, i = _read_varint(b, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = _read_varint(b, i)
        elif w == 2:
            n, i = _read_varint(b, i)
            v = b[i:i + n]
            i += n
        elif w == 1:
            v = b[i:i + 8]
            i += 8
        elif w == 5:
            v = b[i:i + 4]
            i += 4
        else:
            raise ValueError("bad wire type")
        out.append((f, w, v))
    return out


def unmarshal_ListAndWatchResponse(b: bytes) -> List[List[str]]:
    devs = []
    for f, w, v in _fields(b):
        if f == 1 and w == 2:
            d = {1: "", 2: ""}
            for f2, w2, v2 in _fields(v):
                if w2 == 2 and f2 in d:
                    d[f2] = v2.decode()
            devs.append([d[1], d[2]])
    return devs


def unmarshal_AllocateResponse(b: bytes) -> List[Dict[str, str]]:
    """Decoded form: one envs dict per container (the only meaningful comparison — map order is free)."""
    out = []
    for f, w, v in _fields(b):
        if f == 1 and w == 2:
            envs: Dict[str, str] = {}
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    kv = {1: b"", 2: b""}
                    for f3, w3, v3 in _fields(v2):
                        kv[f3] = v3
                    envs[kv[1].decode()] = kv[2].decode()
                else:
                    raise AssertionError("reference never sets mounts/devices/annotations")
            out.append(envs)
    return out


# ---------------------------------------------------------------- health (nvidia.go:100-152, server.go:172-189)


def xid_event_effects(dev_ids: List[str], etype: int, edata: int, uuid: Optional[str],
                      wait_error: bool = False) -> List[int]:
    """Indices of fake devices that watchXIDs pushes on the xids channel for one WaitForEvent result."""
    XidCriticalError = 0x8  # nvml.h:1082
    if wait_error and etype != XidCriticalError:  # nvidia.go:127-129
        return []
    if edata in (31, 43, 45):  # nvidia.go:134-136
        return []
    if uuid is None or len(uuid) == 0:  # nvidia.go:138-144
        return list(range(len(dev_ids)))
    return [i for i, d in enumerate(dev_ids) if extractRealDeviceID(d) == uuid]  # nvidia.go:146-150


def list_and_watch_stream(devs: List[List[str]], unhealthy_events: List[int]) -> List[bytes]:
    """server.go:172-185: the full list once, then the full list again after EVERY health event
    (one event per fake device; Unhealthy is sticky)."""
    devs = [list(d) for d in devs]
    frames = [marshal_ListAndWatchResponse(devs)]
    for i in unhealthy_events:
        devs[i][1] = Unhealthy
        frames.append(marshal_ListAndWatchResponse(devs))
    return frames


# ---------------------------------------------------------------- podutils.go / podmanager.go


def quantity_value(q) -> int:
    """resource.Quantity.Value() for the forms a gpu-mem limit takes: integers, optionally with a
    decimal-SI or binary-SI suffix; fractional values round up. (Value() rounds AWAY FROM ZERO; this restatement — and
    the product's — rounds up, which differs only for negative fractional quantities; the apiserver refuses negative
    resource limits at admission, so no pod the plugin can see carries one.)"""
    if isinstance(q, int):
        return q
    s = str(q).strip()
    suffixes = {"Ki": 1 << 10, "Mi": 1 << 20, "Gi": 1 << 30, "Ti": 1 << 40, "Pi": 1 << 50, "Ei": 1 << 60,
                "k": 10 ** 3, "M": 10 ** 6, "G": 10 ** 9, "T": 10 ** 12, "P": 10 ** 15, "E": 10 ** 18}
    mult_num, mult_den = 1, 1
    for suf, m in suffixes.items():
        if s.endswith(suf):
            s, mult_num = s[: -len(suf)], m
            break
    else:
        if s.endswith("m"):
            s, mult_den = s[:-1], 1000
    from fractions import Fraction
    import math
    return math.ceil(Fraction(s) * mult_num / mult_den)


def getGPUMemoryFromPodResource(pod: dict) -> int:  # podutils.go:122-131
    total = 0
    for c in pod.get("spec", {}).get("containers", []) or []:
        limits = (c.get("resources") or {}).get("limits") or {}
        if resourceName in limits:
            total += quantity_value(limits[resourceName])
    return total


def _atoi(s: str) -> Optional[int]:  # strconv.Atoi
    t = s[1:] if s[:1] in "+-" else s
    if not t or not all("0" <= ch <= "9" for ch in t):
        return None
    v = int(s)
    if not -(1 << 63) <= v < (1 << 63):
        return None
    return v


def getGPUIDFromPodAnnotation(pod: dict) -> int:  # podutils.go:37-61
    ann = (pod.get("metadata") or {}).get("annotations") or {}
    id_ = -1
    if len(ann) > 0 and EnvResourceIndex in ann:
        v = _atoi(ann[EnvResourceIndex])
        id_ = v if v is not None else -1
    return id_


def getAssumeTimeFromPodAnnotation(pod: dict) -> int:  # podutils.go:64-75
    ann = (pod.get("metadata") or {}).get("annotations") or {}
    s = ann.get(EnvResourceAssumeTime)
    if s is not None and s != "" and all("0" <= ch <= "9" for ch in s) and int(s) < (1 << 64):
        return int(s)
    return 0


def isGPUMemoryAssumedPod(pod: dict) -> bool:  # podutils.go:78-119
    if getGPUMemoryFromPodResource(pod) <= 0:
        return False
    ann = (pod.get("metadata") or {}).get("annotations") or {}
    if EnvResourceAssumeTime not in ann:
        return False
    return ann.get(EnvAssignedFlag) == "false" if EnvAssignedFlag in ann else False


def go110_sort(keys: List[int], items: list) -> list:
    """sort.Sort(orderedPodByAssumeTime) as Go 1.10 runs it (the reference builds with golang:1.10, Dockerfile:1,
    .travis.yml:3-4), with the reference's NON-STRICT Less (`<=`, podmanager.go:256-258). The order of pods with EQUAL
    assume-times is whatever this algorithm leaves behind, so it is restated whole: quickSort (slices of <= 12
    elements: one ShellSort pass with gap 6, then insertionSort), doPivot (median of three, Tukey's ninther above 40
    elements, the duplicate-protection pass), heapSort once 2*ceil(lg(n+1)) levels are used up (which an all-tied
    input reaches: with `<=` every partition is maximally lopsided). Distinct keys sort identically under any
    algorithm.

    Third-party dependency absent from /root/reference: Go standard library, package sort, go1.10 (src/sort/sort.go:
    insertionSort, siftDown, heapSort, medianOfThree, doPivot, quickSort, Sort, maxDepth), restated from its
    published source. PARITY UNPINNED: no Go toolchain here to run it."""
    data = list(range(len(items)))  # data[i] = index into items; Less/Swap act on positions

    def less(i, j):
        return keys[data[i]] <= keys[data[j]]

    def swap(i, j):
        data[i], data[j] = data[j], data[i]

    def insertion_sort(a, b):
        for i in range(a + 1, b):
            j = i
            while j > a and less(j, j - 1):
                swap(j, j - 1)
                j -= 1

    def sift_down(lo, hi, first):
        root = lo
        while True:
            child = 2 * root + 1
            if child >= hi:
                return
            if child + 1 < hi and less(first + child, first + child + 1):
                child += 1
            if not less(first + root, first + child):
                return
            swap(first + root, first + child)
            root = child

    def heap_sort(a, b):
        first, lo, hi = a, 0, b - a
        for i in range((hi - 1) // 2, -1, -1):
            sift_down(i, hi, first)
        for i in range(hi - 1, -1, -1):
            swap(first, first + i)
            sift_down(lo, i, first)

    def median_of_three(m1, m0, m2):
        if less(m1, m0):
            swap(m1, m0)
        if less(m2, m1):
            swap(m2, m1)
            if less(m1, m0):
                swap(m1, m0)

    def do_pivot(lo, hi):
        m = (lo + hi) >> 1
        if hi - lo > 40:
            s_ = (hi - lo) // 8
            median_of_three(lo, lo + s_, lo + 2 * s_)
            median_of_three(m, m - s_, m + s_)
            median_of_three(hi - 1, hi - 1 - s_, hi - 1 - 2 * s_)
        median_of_three(lo, m, hi - 1)
        pivot = lo
        a, c = lo + 1, hi - 1
        while a < c and less(a, pivot):
            a += 1
        b = a
        while True:
            while b < c and not less(pivot, b):
                b += 1
            while b < c and less(pivot, c - 1):
                c -= 1
            if b >= c:
                break
            swap(b, c - 1)
            b += 1
            c -= 1
        protect = hi - c < 5
        if not protect and hi - c < (hi - lo) // 4:
            dups = 0
            if not less(pivot, hi - 1):
                swap(c, hi - 1)
                c += 1
                dups += 1
            if not less(b - 1, pivot):
                b -= 1
                dups += 1
            if not less(m, pivot):
                swap(m, b - 1)
                b -= 1
                dups += 1
            protect = dups > 1
        if protect:
            while True:
                while a < b and not less(b - 1, pivot):
                    b -= 1
                while a < b and less(a, pivot):
                    a += 1
                if a >= b:
                    break
                swap(a, b - 1)
                a += 1
                b -= 1
        swap(pivot, b - 1)
        return b - 1, c

    def quick_sort(a, b, max_depth):
        while b - a > 12:
            if max_depth == 0:
                heap_sort(a, b)
                return
            max_depth -= 1
            mlo, mhi = do_pivot(a, b)
            if mlo - a < b - mhi:
                quick_sort(a, mlo, max_depth)
                a = mhi
            else:
                quick_sort(mhi, b, max_depth)
                b = mlo
        if b - a > 1:
            for i in range(a + 6, b):
                if less(i, i - 6):
                    swap(i, i - 6)
            insertion_sort(a, b)

    n = len(data)
    depth, i = 0, n
    while i > 0:
        depth += 1
        i >>= 1
    quick_sort(0, n, depth * 2)
    return [items[i] for i in data]


def getCandidatePods(pod_list: List[dict], nodeName: str) -> List[dict]:  # podmanager.go:162-262
    pods, seen = [], set()
    for pod in pod_list:
        if (pod.get("spec") or {}).get("nodeName") != nodeName:
            continue
        uid = (pod.get("metadata") or {}).get("uid")
        if uid not in seen:
            pods.append(pod)
            seen.add(uid)
    cand = [p for p in pods if isGPUMemoryAssumedPod(p)]
    return go110_sort([getAssumeTimeFromPodAnnotation(p) for p in cand], cand)


def patchPodAnnotationSpecAssigned(now_unix_nano: int) -> bytes:  # podutils.go:27-35 (json.Marshal sorts keys)
    return json.dumps({"metadata": {"annotations": {EnvAssignedFlag: "true",
                                                    EnvResourceAssumeTime: "%d" % now_unix_nano}}},
                      separators=(",", ":"), sort_keys=True).encode()


# ---------------------------------------------------------------- allocate.go


def buildErrResponse(container_requests: List[List[str]], podReqGPU: int, metric: str, gpuMemory: int):
    return [{envNVGPU: "no-gpu-has-%d%s-to-run" % (podReqGPU, metric),  # allocate.go:24-39
             EnvResourceIndex: "-1",
             EnvResourceByPod: "%d" % podReqGPU,
             EnvResourceByContainer: "%d" % len(req),
             EnvResourceByDev: "%d" % gpuMemory} for req in container_requests]


def Allocate(container_requests: List[List[str]], pending_pods: List[dict], nodeName: str,
             devNameMap: Dict[str, int], gpuMemory: int, metric: str = GiBPrefix,
             disableCGPUIsolation: bool = False, list_error: bool = False, patch=None):
    """allocate.go:42-198. Returns (container env dicts, pod that was patched or None).
    `patch(pod, body_bytes) -> Optional[str]` performs the annotation PATCH and returns an error
    string or None; it is retried once iff the error equals OptimisticLockErrorMsg (:135-149)."""
    podReqGPU = sum(len(r) for r in container_requests)  # :54-56
    if list_error:  # :62-66
        return buildErrResponse(container_requests, podReqGPU, metric, gpuMemory), None
    pods = getCandidatePods(pending_pods, nodeName)
    assumePod = None
    for pod in pods:  # :78-88
        if getGPUMemoryFromPodResource(pod) == podReqGPU:
            assumePod = pod
            break
    if assumePod is not None:
        id_ = getGPUIDFromPodAnnotation(assumePod)  # :91
        if id_ >= 0:
            devIndxMap = {v: k for k, v in devNameMap.items()}  # server.go:72-83
            if id_ not in devIndxMap:
                id_ = -1
        if id_ < 0:
            return buildErrResponse(container_requests, podReqGPU, metric, gpuMemory), None  # :108-110
        responses = []
        for req in container_requests:  # :113-128
            envs = {envNVGPU: "%d" % id_,
                    EnvResourceIndex: "%d" % id_,
                    EnvResourceByPod: "%d" % podReqGPU,
                    EnvResourceByContainer: "%d" % len(req),
                    EnvResourceByDev: "%d" % gpuMemory}
            if disableCGPUIsolation:
                envs["CGPU_DISABLE"] = "true"
            responses.append(envs)
        if patch is not None:  # :130-149
            import time
            body = patchPodAnnotationSpecAssigned(time.time_ns())
            err = patch(assumePod, body)
            if err is not None:
                if err == OptimisticLockErrorMsg:
                    err = patch(assumePod, body)
                    if err is not None:
                        return buildErrResponse(container_requests, podReqGPU, metric, gpuMemory), None
                else:
                    return buildErrResponse(container_requests, podReqGPU, metric, gpuMemory), None
        return responses, assumePod
    if len(devNameMap) == 1:  # :151-177
        (devName, devIndex), = devNameMap.items()
        responses = []
        for req in container_requests:
            envs = {envNVGPU: devName,
                    EnvResourceIndex: "%d" % devIndex,
                    EnvResourceByPod: "%d" % podReqGPU,
                    EnvResourceByContainer: "%d" % len(req),
                    EnvResourceByDev: "%d" % gpuMemory}
            if disableCGPUIsolation:
                envs["CGPU_DISABLE"] = "true"
            responses.append(envs)
        return responses, None
    return buildErrResponse(container_requests, podReqGPU, metric, gpuMemory), None  # :179-184
