"""Allocate on CPU: gsb_allocate (C ABI: wire bytes in, wire bytes out) against the pure-Python
restatement of allocate.go (oracle/wire_oracle.Allocate), on known answers, SURVEY §8(d) config 4,
tie/edge cases and randomized clusters. Responses are compared DECODED (gogo emits map entries in
Go-map order, so only the decoded envs are defined by the reference)."""
import ctypes as C
import json
import os

import pytest
from hypothesis import given, settings, strategies as st

from gpushare_device_plugin_b200 import _abi
from gpushare_device_plugin_b200.nvidia.allocate import AllocateContext, pod_table
from gpushare_device_plugin_b200.testing.mock_kube import config4_pods, make_pod
from oracle import wire_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "wire_kat.json")))
NODE = "b200-0"
UUIDS = ["GPU-%08x-4820-abfc-e83e-9431819757%02x" % (0xfef80890 + i, i) for i in range(8)]
MINORS8 = {u: m for u, m in zip(UUIDS, [2, 3, 0, 1, 6, 7, 4, 5])}  # minor != index, as on real HGX boxes


def ids(n, g=0):
    return [wo.generateFakeDeviceID(UUIDS[g], j) for j in range(n)]


def product_allocate(container_requests, pods, devNameMap, slices=179, unit_gib=True, cgpu=False, node=NODE,
                     pods_unique=False):
    actx = AllocateContext(devNameMap, slices, unit_gib, cgpu)
    actx.ctx.pods_unique = 1 if pods_unique else 0
    req = wo.marshal_AllocateRequest(container_requests)
    table, _keep = pod_table(pods, node)
    buf = C.create_string_buffer(1 << 16)
    n, pidx, preq = C.c_size_t(0), C.c_int32(-1), C.c_uint32(0)
    kind = _abi.lib.gsb_allocate(C.byref(actx.ctx), table, len(pods), req, len(req), buf, len(buf), C.byref(n),
                                 C.byref(pidx), C.byref(preq))
    assert kind > 0, _abi.last_error()
    return kind, wo.unmarshal_AllocateResponse(buf.raw[: n.value]), pidx.value, preq.value, buf.raw[: n.value]


def both(container_requests, pods, devNameMap, **kw):
    kind, envs, pidx, preq, raw = product_allocate(container_requests, pods, devNameMap, **kw)
    want, want_pod = wo.Allocate(container_requests, pods, kw.get("node", NODE), devNameMap, kw.get("slices", 179),
                                 wo.GiBPrefix if kw.get("unit_gib", True) else wo.MiBPrefix, kw.get("cgpu", False))
    assert envs == want
    if want_pod is None:
        assert pidx == -1
    else:
        assert pods[pidx]["metadata"]["uid"] == want_pod["metadata"]["uid"]
    assert preq == sum(len(r) for r in container_requests)
    # our entry order is sorted keys: one of the orders gogo may produce
    assert raw == wo.marshal_AllocateResponse(want)
    return kind, envs, pidx


def test_kat_error_ok_and_single_gpu_responses():
    k = KAT["err_response"]
    kind, envs, _ = both([ids(4)], [], MINORS8)
    assert kind == _abi.GSB_ALLOC_ERR_RESPONSE and envs == [k["envs"]]
    k = KAT["ok_response"]
    pod = make_pod(0, NODE, gpu_mem=4, idx=3, assume_time=5)
    kind, envs, pidx = both([ids(4)], [pod], MINORS8)
    assert kind == _abi.GSB_ALLOC_MATCHED and envs == [k["envs"]] and pidx == 0
    k = KAT["single_gpu_response"]
    kind, envs, _ = both([ids(2)], [], {k["uuid"]: k["minor"]})
    assert kind == _abi.GSB_ALLOC_SINGLE_GPU and envs == [k["envs"]]


def test_config4_binpack_64_pods():
    """64 pods @4 GiB, IDX = i div 8, assume time ascending: the i-th Allocate answers IDX i div 8; the
    PATCH (simulated here by flipping the annotation) removes the pod from the candidates."""
    pods = config4_pods(NODE)
    minors = {u: i for i, u in enumerate(UUIDS)}
    for i in range(64):
        kind, envs, pidx = both([ids(4)], pods, minors)
        assert kind == _abi.GSB_ALLOC_MATCHED and pidx == i
        assert envs == [{"NVIDIA_VISIBLE_DEVICES": str(i // 8), "ALIYUN_COM_GPU_MEM_IDX": str(i // 8),
                         "ALIYUN_COM_GPU_MEM_POD": "4", "ALIYUN_COM_GPU_MEM_CONTAINER": "4",
                         "ALIYUN_COM_GPU_MEM_DEV": "179"}]
        pods[pidx]["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] = "true"
    kind, _, _ = both([ids(4)], pods, minors)
    assert kind == _abi.GSB_ALLOC_ERR_RESPONSE  # nothing left to assign


@pytest.mark.parametrize("k", KAT["candidate_predicate"], ids=lambda k: f"{k['limit']}-{k['assume']}-{k['assigned']}")
def test_candidate_predicate_truth_table(k):
    pod = make_pod(0, NODE, gpu_mem=max(k["limit"], 1), idx=1, assume_time=None, assigned=k["assigned"])
    if k["limit"] == 0:
        pod["spec"]["containers"][0]["resources"]["limits"] = {}
    if k["assume"] is not None:
        pod["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSUME_TIME"] = k["assume"]
    assert wo.isGPUMemoryAssumedPod(pod) == k["candidate"]
    kind, _, pidx = both([ids(max(k["limit"], 1))], [pod], MINORS8)
    assert (kind == _abi.GSB_ALLOC_MATCHED) == k["candidate"]


def test_multi_container_cgpu_mib_and_idx_edge_cases():
    pod = make_pod(1, NODE, gpu_mem=6, idx=7, assume_time=9, containers=2)
    kind, envs, _ = both([ids(2), ids(4, 1)], [pod], MINORS8, cgpu=True)
    assert kind == _abi.GSB_ALLOC_MATCHED and [e["ALIYUN_COM_GPU_MEM_CONTAINER"] for e in envs] == ["2", "4"]
    assert all(e["CGPU_DISABLE"] == "true" and e["ALIYUN_COM_GPU_MEM_POD"] == "6" for e in envs)
    # MiB unit changes only the error text
    kind, envs, _ = both([ids(3)], [], MINORS8, unit_gib=False, slices=183359)
    assert envs[0]["NVIDIA_VISIBLE_DEVICES"] == "no-gpu-has-3MiB-to-run" and envs[0]["ALIYUN_COM_GPU_MEM_DEV"] == "183359"
    # IDX missing / unparsable / not a minor on this node -> error response, pod NOT claimed
    for idx in (None, "x7", "8", "-2", "+3"):
        p = make_pod(2, NODE, gpu_mem=4, idx=None, assume_time=1)
        if idx is not None:
            p["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_IDX"] = idx
        kind, envs, pidx = both([ids(4)], [p], MINORS8)
        if idx == "+3":  # strconv.Atoi accepts a sign
            assert kind == _abi.GSB_ALLOC_MATCHED and envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "3"
        else:
            assert kind == _abi.GSB_ALLOC_ERR_RESPONSE and pidx == -1
    # empty request: no containers, empty response
    kind, envs, _ = both([], [], MINORS8)
    assert envs == []
    # pods of another node and duplicate UIDs are ignored
    other = make_pod(3, "b200-1", gpu_mem=4, idx=0, assume_time=1)
    dup_a, dup_b = make_pod(4, NODE, gpu_mem=4, idx=1, assume_time=50), make_pod(4, NODE, gpu_mem=4, idx=2, assume_time=10)
    kind, envs, pidx = both([ids(4)], [other, dup_a, dup_b], MINORS8)
    assert pidx == 1 and envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "1"


def test_assume_time_order_and_go110_tie_rule():
    # oldest assume time wins among pods with the same limit
    pods = [make_pod(i, NODE, gpu_mem=4, idx=i % 8, assume_time=t) for i, t in enumerate([30, 10, 20, 10 ** 19, 5])]
    _, envs, pidx = both([ids(4)], pods, {u: i for i, u in enumerate(UUIDS)})
    assert pidx == 4
    # ties: Go 1.10 sort.Sort with Less `<=` — insertion pass moves the LATER equal element first
    ties = [make_pod(i, NODE, gpu_mem=4, idx=i % 8, assume_time=7) for i in range(5)]
    _, _, pidx = both([ids(4)], ties, {u: i for i, u in enumerate(UUIDS)})
    assert pidx == 4
    # 7..12 elements: the gap-6 ShellSort pass runs first
    ties = [make_pod(i, NODE, gpu_mem=4, idx=i % 8, assume_time=7) for i in range(9)]
    both([ids(4)], ties, {u: i for i, u in enumerate(UUIDS)})
    # unparsable / overflowing assume time counts as 0 (oldest)
    pods = [make_pod(0, NODE, gpu_mem=4, idx=1, assume_time=3), make_pod(1, NODE, gpu_mem=4, idx=2, assume_time=None)]
    pods[1]["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSUME_TIME"] = str(1 << 64)
    _, _, pidx = both([ids(4)], pods, MINORS8)
    assert pidx == 1


pod_strategy = st.builds(
    lambda i, mem, idx, t, assigned, node, has_t: _mk(i, mem, idx, t, assigned, node, has_t),
    st.integers(0, 40), st.integers(0, 6), st.sampled_from([None, "0", "1", "5", "7", "9", "-1", "abc"]),
    st.integers(0, 6), st.sampled_from([None, "false", "true", "False"]), st.sampled_from([NODE, NODE, NODE, "other"]),
    st.booleans())


def _mk(i, mem, idx, t, assigned, node, has_t):
    p = make_pod(i, node, gpu_mem=max(mem, 1), idx=None, assume_time=t if has_t else None, assigned=assigned)
    if mem == 0:
        p["spec"]["containers"][0]["resources"]["limits"] = {}
    if idx is not None:
        p["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_IDX"] = idx
    return p


@settings(max_examples=300, deadline=None)
@given(pods=st.lists(pod_strategy, max_size=14), reqs=st.lists(st.integers(0, 4), max_size=3), one_gpu=st.booleans(),
       cgpu=st.booleans())
def test_randomized_clusters_match_oracle(pods, reqs, one_gpu, cgpu):
    dev = {UUIDS[0]: 5} if one_gpu else MINORS8
    both([ids(n) for n in reqs], pods, dev, cgpu=cgpu)


@settings(max_examples=150, deadline=None)
@given(pods=st.lists(pod_strategy, min_size=13, max_size=60), reqs=st.lists(st.integers(0, 4), max_size=3),
       one_gpu=st.booleans())
def test_large_clusters_match_oracle(pods, reqs, one_gpu):
    """> 12 candidates: the one-pass selection (no sort) against the oracle's full sort, duplicate UIDs and
    tied assume-times included."""
    both([ids(n) for n in reqs], pods, {UUIDS[0]: 5} if one_gpu else MINORS8)


@settings(max_examples=150, deadline=None)
@given(pods=st.lists(pod_strategy, max_size=40, unique_by=lambda p: p["metadata"]["uid"]),
       reqs=st.lists(st.integers(0, 4), max_size=3))
def test_pods_unique_hint_changes_nothing_on_a_uid_keyed_table(pods, reqs):
    """gsb_allocate_ctx.pods_unique skips podmanager.go's dedupe; on a table without repeated UIDs the decision and
    the response bytes are the same either way."""
    cr = [ids(n) for n in reqs]
    a = product_allocate(cr, pods, MINORS8)
    b = product_allocate(cr, pods, MINORS8, pods_unique=True)
    assert a == b
    both(cr, pods, MINORS8)


def test_malformed_request_is_rejected_not_guessed():
    actx = AllocateContext(MINORS8, 179, True, False)
    buf = C.create_string_buffer(1024)
    n, pidx, preq = C.c_size_t(0), C.c_int32(-1), C.c_uint32(0)
    for bad in (b"\x0a\x05\x0a", b"\x0a\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01", b"\x0f"):
        rc = _abi.lib.gsb_allocate(C.byref(actx.ctx), None, 0, bad, len(bad), buf, len(buf), C.byref(n),
                                   C.byref(pidx), C.byref(preq))
        assert rc == _abi.GSB_ERR_MALFORMED
    # unknown fields are skipped like any proto3 reader does
    ok = b"\x10\x05" + wo.marshal_AllocateRequest([ids(2)])
    rc = _abi.lib.gsb_allocate(C.byref(actx.ctx), None, 0, ok, len(ok), buf, len(buf), C.byref(n), C.byref(pidx),
                               C.byref(preq))
    assert rc == _abi.GSB_ALLOC_ERR_RESPONSE and preq.value == 2


# ---- request decoding: same accept / reject as gogo's generated Unmarshal (api.pb.go:2141-2300, 2991-3089) ----------

def _product_decode(req: bytes):
    """(accepted, devicesIDs-per-container) as gsb_allocate sees the request: on a one-GPU node every decodable
    request is answered, with POD = the total and CONTAINER = the per-container count."""
    actx = AllocateContext({UUIDS[0]: 0}, 179, True, False)
    buf = C.create_string_buffer(1 << 20)
    n, pidx, preq = C.c_size_t(0), C.c_int32(-1), C.c_uint32(0)
    rc = _abi.lib.gsb_allocate(C.byref(actx.ctx), None, 0, req, len(req), buf, len(buf), C.byref(n), C.byref(pidx),
                               C.byref(preq))
    if rc == _abi.GSB_ERR_MALFORMED:
        return False, None
    assert rc > 0, rc
    return True, [int(e["ALIYUN_COM_GPU_MEM_CONTAINER"]) for e in wo.unmarshal_AllocateResponse(buf.raw[: n.value])]


def _oracle_decode(req: bytes):
    try:
        return True, [len(ids) for ids in wo.unmarshal_AllocateRequest(req)]
    except wo.UnmarshalError:
        return False, None


_valid_requests = st.lists(st.lists(st.text(alphabet="GPU-0123456789abcdef_", max_size=50), max_size=5), max_size=4) \
    .map(wo.marshal_AllocateRequest)
_unknown_field = st.one_of(
    st.builds(lambda f, v: wo.encodeVarintApi(f << 3 | 0) + wo.encodeVarintApi(v), st.integers(2, 40), st.integers(0, 1 << 40)),
    st.builds(lambda f, b: wo.encodeVarintApi(f << 3 | 2) + wo.encodeVarintApi(len(b)) + b, st.integers(2, 40), st.binary(max_size=12)),
    st.builds(lambda f, b: wo.encodeVarintApi(f << 3 | 1) + b, st.integers(2, 40), st.binary(min_size=8, max_size=8)),
    st.builds(lambda f, b: wo.encodeVarintApi(f << 3 | 5) + b, st.integers(2, 40), st.binary(min_size=4, max_size=4)),
    st.builds(lambda f, g: wo.encodeVarintApi(f << 3 | 3) + g + wo.encodeVarintApi(f << 3 | 4), st.integers(2, 40),
              st.sampled_from([b"", b"\x10\x01", b"\x1a\x02ab", b"\x23\x24", b"\x23\x10\x05\x24"])))


@settings(max_examples=400, deadline=None)
@given(parts=st.lists(st.one_of(_valid_requests, _unknown_field, st.binary(max_size=6)), max_size=5),
       cut=st.integers(0, 400), flip=st.lists(st.tuples(st.integers(0, 400), st.integers(0, 255)), max_size=3))
def test_request_decoder_accepts_and_rejects_what_gogo_does(parts, cut, flip):
    req = bytearray(b"".join(parts))
    for pos, val in flip:  # point mutations: wrong wire types, broken lengths, stray end-groups, tag 0 ...
        if req:
            req[pos % len(req)] = val
    req = bytes(req[: max(0, len(req) - cut % 7)]) if cut % 3 == 0 else bytes(req)
    assert _product_decode(req) == _oracle_decode(req), req.hex()


def test_request_decoder_known_edge_cases():
    cases = {
        b"": (True, []),
        b"\x0a\x00": (True, [0]),
        b"\x0a\x03\x0a\x01a\x0a\x00": (True, [1, 0]),
        b"\x08\x01": (False, None),                    # field 1 as varint: "wrong wireType"
        b"\x0c": (False, None),                        # stray end-group
        b"\x00\x00": (False, None),                    # tag 0
        b"\x0a\x02\x08\x01": (False, None),            # devicesIDs as varint
        b"\x13\x08\x01\x14": (True, []),              # unknown group field 2 skipped
        b"\x13\x08\x01": (False, None),                # ... unterminated
        b"\x16": (False, None),                        # wire type 6
        b"\x10" + b"\xff" * 10 + b"\x01": (False, None),  # 11-byte varint: integer overflow
        b"\x10" + b"\xff" * 9 + b"\x7f": (True, []),     # 10-byte varint is fine
        b"\x12\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01": (False, None),  # length with bit 63 set: negative
        b"\x89\x80\x80\x80\x80\x01\x00": (False, None),  # tag (1<<32|1)<<3|1: int32 wraps to field 1, wire type 1: "wrong wireType"
        b"\x8a\x80\x80\x80\x80\x01\x00": (True, [0]),   # same wrap with wire type 2: read as container_requests
    }
    for req, want in cases.items():
        assert _oracle_decode(req) == want, req.hex()
        assert _product_decode(req) == want, req.hex()


# ---- google.protobuf as an independent witness of the Allocate wire bytes ----------------------------------------

from .test_wire import pb  # noqa: E402,F401  (module-scoped fixture: descriptors with the reference's field numbers)


@settings(max_examples=100, deadline=None)
@given(reqs=st.lists(st.lists(st.text(alphabet="GPU-0123456789abcdef_", min_size=1, max_size=48), max_size=6), max_size=4),
       pods=st.lists(pod_strategy, max_size=10), one_gpu=st.booleans())
def test_protobuf_library_reads_and_writes_the_same_allocate_bytes(pb, reqs, pods, one_gpu):  # noqa: F811
    """Requests serialised by google.protobuf are decoded by gsb_allocate like the oracle's; the response bytes parse
    with google.protobuf into the envs the oracle computes (map entries, repeated messages, field numbers)."""
    m = pb["AllocateRequest"]()
    for ids in reqs:
        m.container_requests.add().devicesIDs.extend(ids)
    req = m.SerializeToString()
    assert req == wo.marshal_AllocateRequest(reqs)
    dev = {UUIDS[0]: 5} if one_gpu else MINORS8
    actx = AllocateContext(dev, 179, True, False)
    table, _keep = pod_table(pods, NODE)
    buf = C.create_string_buffer(1 << 16)
    n, pidx, preq = C.c_size_t(0), C.c_int32(-1), C.c_uint32(0)
    assert _abi.lib.gsb_allocate(C.byref(actx.ctx), table, len(pods), req, len(req), buf, len(buf), C.byref(n),
                                 C.byref(pidx), C.byref(preq)) > 0
    got = pb["AllocateResponse"].FromString(buf.raw[: n.value])
    want, _ = wo.Allocate(reqs, pods, NODE, dev, 179, wo.GiBPrefix, False)
    assert [dict(c.envs) for c in got.container_responses] == want
