"""Host half on CPU: the C ABI's slice arithmetic, IDs and wire encoders against (1) hand-derived
known-answer vectors, (2) the pure-Python restatement of the reference (oracle/wire_oracle.py),
(3) google.protobuf's own encoder/decoder driven by a descriptor with the reference's field numbers
(vendor/.../v1beta1/api.proto:27-161)."""
import json
import os

import pytest
from hypothesis import given, settings, strategies as st

from gpushare_device_plugin_b200 import _abi, device
from oracle import wire_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "wire_kat.json")))
UUID = KAT["fake_id"]["uuid"]


def uuids(n):
    return ["GPU-%08x-4820-abfc-e83e-9431819757%02x" % (0xfef80890 + i, i) for i in range(n)]


# ---- protobuf-library witness -----------------------------------------------------------------

@pytest.fixture(scope="module")
def pb():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="api.proto", package="v1beta1", syntax="proto3")
    S = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = tname
        return m
    msg("Device", [("ID", 1, S.TYPE_STRING, S.LABEL_OPTIONAL, None), ("health", 2, S.TYPE_STRING, S.LABEL_OPTIONAL, None)])
    msg("ListAndWatchResponse", [("devices", 1, S.TYPE_MESSAGE, S.LABEL_REPEATED, ".v1beta1.Device")])
    msg("RegisterRequest", [("version", 1, S.TYPE_STRING, S.LABEL_OPTIONAL, None),
                            ("endpoint", 2, S.TYPE_STRING, S.LABEL_OPTIONAL, None),
                            ("resource_name", 3, S.TYPE_STRING, S.LABEL_OPTIONAL, None)])
    msg("ContainerAllocateRequest", [("devicesIDs", 1, S.TYPE_STRING, S.LABEL_REPEATED, None)])
    msg("AllocateRequest", [("container_requests", 1, S.TYPE_MESSAGE, S.LABEL_REPEATED, ".v1beta1.ContainerAllocateRequest")])
    car = msg("ContainerAllocateResponse", [("envs", 1, S.TYPE_MESSAGE, S.LABEL_REPEATED, ".v1beta1.ContainerAllocateResponse.EnvsEntry")])
    e = car.nested_type.add(name="EnvsEntry")
    e.field.add(name="key", number=1, type=S.TYPE_STRING, label=S.LABEL_OPTIONAL)
    e.field.add(name="value", number=2, type=S.TYPE_STRING, label=S.LABEL_OPTIONAL)
    e.options.map_entry = True
    msg("AllocateResponse", [("container_responses", 1, S.TYPE_MESSAGE, S.LABEL_REPEATED, ".v1beta1.ContainerAllocateResponse")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("v1beta1." + n))  # noqa: E731
    return {n: get(n) for n in ("Device", "ListAndWatchResponse", "RegisterRequest", "AllocateRequest", "AllocateResponse")}


# ---- known answers ---------------------------------------------------------------------------------

def test_fake_device_ids_kat():
    k = KAT["fake_id"]
    assert device.fake_device_id(k["uuid"], k["j"]) == k["id"] == wo.generateFakeDeviceID(k["uuid"], k["j"])
    assert device.real_device_id(k["id"]) == k["uuid"] == wo.extractRealDeviceID(k["id"])
    assert device.real_device_id("no-separator") == "no-separator" == wo.extractRealDeviceID("no-separator")
    assert device.real_device_id("a-_-1-_-2") == "a"


@pytest.mark.parametrize("k", KAT["slices"], ids=lambda k: str(k["total_bytes"]))
def test_slice_arithmetic_kat(k):
    mib = wo.mib_from_bytes(k["total_bytes"])
    assert mib == k["mib"]
    assert device.slices(mib, True) == k["gib_slices"] == wo.setGPUMemory(mib, wo.GiBPrefix)
    assert device.slices(mib, False) == k["mib_slices"] == wo.setGPUMemory(mib, wo.MiBPrefix)


@given(st.integers(min_value=0, max_value=(1 << 45)))
def test_double_floor_identity(total_bytes):
    mib = total_bytes // (1 << 20)
    assert device.slices(mib, True) == total_bytes >> 30 == wo.setGPUMemory(wo.mib_from_bytes(total_bytes), wo.GiBPrefix)


@pytest.mark.parametrize("k", KAT["list_and_watch_sizes"], ids=lambda k: f"{k['n_gpus']}x{k['slices']}")
def test_list_and_watch_sizes_kat(k):
    b = device.encode_list_and_watch(uuids(k["n_gpus"]), k["slices"])
    assert len(b) == k["bytes"]
    devs = [[wo.generateFakeDeviceID(u, j), wo.Healthy] for u in uuids(k["n_gpus"]) for j in range(k["slices"])]
    assert b == wo.marshal_ListAndWatchResponse(devs)


def test_list_and_watch_bytes_kat(pb):
    b = device.encode_list_and_watch([UUID], 179)
    first = bytes.fromhex(KAT["list_and_watch_first_device_hex"])
    assert b[: len(first)] == first
    bits = bytearray(23)
    bits[0] = 1
    ub = device.encode_list_and_watch([UUID], 179, bytes(bits))
    bad = bytes.fromhex(KAT["unhealthy_device_hex"])
    assert ub[: len(bad)] == bad and ub[len(bad):] == b[len(first):]
    m = pb["ListAndWatchResponse"]()
    m.ParseFromString(ub)
    assert len(m.devices) == 179 and m.devices[0].health == "Unhealthy" and m.devices[178].ID == UUID + "-_-178"
    assert m.SerializeToString() == ub  # protobuf's own encoder produces the same bytes


def test_register_request_kat(pb):
    b = device.encode_register_request("v1beta1", "aliyungpushare.sock", "aliyun.com/gpu-mem")
    assert b.hex() == KAT["register_request_hex"]
    assert b == wo.marshal_RegisterRequest(wo.Version, wo.serverSockName, wo.resourceName)
    m = pb["RegisterRequest"](version="v1beta1", endpoint="aliyungpushare.sock", resource_name="aliyun.com/gpu-mem")
    assert m.SerializeToString() == b
    assert device.encode_register_request("", "x", "") == b"\x12\x01x"  # proto3 omits empty strings


def test_patch_body_kat():
    import ctypes as C
    buf = C.create_string_buffer(256)
    n = _abi.lib.gsb_patch_assigned_body(KAT["patch_body"]["now_ns"], buf, len(buf))
    assert buf.raw[:n].decode() == KAT["patch_body"]["body"]
    assert wo.patchPodAnnotationSpecAssigned(KAT["patch_body"]["now_ns"]).decode() == KAT["patch_body"]["body"]
    assert _abi.lib.gsb_patch_assigned_body(1, buf, 10) == _abi.GSB_ERR_BUFFER_TOO_SMALL


@pytest.mark.parametrize("k", KAT["xid_table"], ids=lambda k: str(k["xid"]))
def test_xid_filter_kat(k):
    assert bool(_abi.lib.gsb_xid_is_benign(k["xid"])) == (not k["unhealthy"])
    ids = [wo.generateFakeDeviceID(u, j) for u in uuids(2) for j in range(3)]
    hit = wo.xid_event_effects(ids, 8, k["xid"], uuids(2)[1])
    assert hit == ([3, 4, 5] if k["unhealthy"] else [])


# ---- properties ---------------------------------------------------------------------------------

@settings(max_examples=60, deadline=None)
@given(n_gpus=st.integers(0, 9), slices=st.integers(0, 200), data=st.data())
def test_encoder_equals_oracle_and_protobuf(pb, n_gpus, slices, data):
    us = uuids(n_gpus)
    total = n_gpus * slices
    bad = set(data.draw(st.lists(st.integers(0, max(total - 1, 0)), max_size=20))) if total else set()
    bits = bytearray((total + 7) // 8)
    for i in bad:
        bits[i >> 3] |= 1 << (i & 7)
    got = device.encode_list_and_watch(us, slices, bytes(bits) if total else None)
    devs = [[wo.generateFakeDeviceID(u, j), wo.Unhealthy if g * slices + j in bad else wo.Healthy]
            for g, u in enumerate(us) for j in range(slices)]
    assert got == wo.marshal_ListAndWatchResponse(devs)
    assert wo.unmarshal_ListAndWatchResponse(got) == devs
    m = pb["ListAndWatchResponse"]()
    m.ParseFromString(got)
    assert [[d.ID, d.health] for d in m.devices] == devs and m.SerializeToString() == got


def test_buffer_too_small_and_bad_arguments():
    import ctypes as C
    arr = (C.c_char_p * 1)(UUID.encode())
    need = _abi.lib.gsb_encode_list_and_watch(arr, 1, 179, None, None, 0)
    assert need == 10451
    buf = C.create_string_buffer(100)
    assert _abi.lib.gsb_encode_list_and_watch(arr, 1, 179, None, buf, 100) == _abi.GSB_ERR_BUFFER_TOO_SMALL
    assert _abi.lib.gsb_encode_list_and_watch(None, 1, 179, None, None, 0) == _abi.GSB_ERR_INVALID_ARGUMENT
    small = C.create_string_buffer(10)
    assert _abi.lib.gsb_fake_device_id(UUID.encode(), 1, small, 10) == _abi.GSB_ERR_BUFFER_TOO_SMALL
    assert all(len(device.fake_device_id(UUID, j)) <= 63 for j in (0, 178, 183358))  # api.proto:82-85 limit


def test_reference_stream_amplification_vs_coalesced():
    """One XID on one of 8 GPUs: the reference re-sends the full list once per fake device (179 frames,
    server.go:172-185); the final frame is what a coalesced single resend carries."""
    us = uuids(8)
    devs = [[wo.generateFakeDeviceID(u, j), wo.Healthy] for u in us for j in range(179)]
    hit = wo.xid_event_effects([d[0] for d in devs], 8, 79, us[5])
    frames = wo.list_and_watch_stream(devs, hit)
    assert len(frames) == 180 and sum(map(len, frames[1:])) == 14_998_052  # ~15 MB for one XID
    bits = bytearray((len(devs) + 7) // 8)
    for i in hit:
        bits[i >> 3] |= 1 << (i & 7)
    assert device.encode_list_and_watch(us, 179, bytes(bits)) == frames[-1]
