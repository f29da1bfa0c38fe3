"""The device-plugin gRPC surface end to end on CPU: real grpc over unix sockets, a fake kubelet
(Registration server + device-manager client), a stateful mock apiserver; inventory injected
(tests/fakes.py). Wire bytes are checked against the oracle's restatement of the reference."""
import json
import os
import queue
import re
import threading
import time

import grpc
import pytest

from gpushare_device_plugin_b200 import device
from gpushare_device_plugin_b200.kubelet.client import KubeletClientConfig, NewKubeletClient
from gpushare_device_plugin_b200.nvidia import const, kubeclient, podmanager, server
from gpushare_device_plugin_b200.testing.fake_kubelet import FakeKubelet
from gpushare_device_plugin_b200.testing.mock_kube import MockKube, config4_pods, make_node, make_pod
from oracle import wire_oracle as wo

from . import fakes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "wire_kat.json")))
NODE = "b200-0"


@pytest.fixture
def world(tmp_path, monkeypatch):
    fakes.install(monkeypatch)
    monkeypatch.setattr(time, "sleep", lambda s: None)  # the retry back-offs of podmanager.go
    kube = MockKube(make_node(NODE, labels={}), config4_pods(NODE))
    podmanager.kubeInit(kubeclient.Clientset(kube.url), NODE)
    kubelet = FakeKubelet(str(tmp_path))
    made = []

    def make(**kw):
        p = server.NewNvidiaDevicePlugin(False, kw.pop("healthCheck", True), kw.pop("queryKubelet", False),
                                         kw.pop("client", None), socket=str(tmp_path / "aliyungpushare.sock"), **kw)
        made.append(p)
        return p
    yield type("W", (), {"kube": kube, "kubelet": kubelet, "make": staticmethod(make), "dir": tmp_path})
    for p in made:
        p.Stop()
    kubelet.stop()
    kube.close()


def all_devs(unhealthy=()):
    return [[wo.generateFakeDeviceID(u, j), wo.Unhealthy if (g, j) in unhealthy or g in unhealthy else wo.Healthy]
            for g, u in enumerate(fakes.UUIDS) for j in range(179)]


class Frames:
    """One pump thread per stream: frames land in a queue (a timed-out next() must not be abandoned
    mid-call, it would swallow the following frame)."""

    def __init__(self, call):
        self.q = queue.Queue()

        def pump():
            try:
                for f in call:
                    self.q.put(f)
            except grpc.RpcError:
                pass
            self.q.put(None)
        threading.Thread(target=pump, daemon=True).start()

    def __iter__(self):
        return self


def next_frame(frames, timeout=5.0):
    try:
        return frames.q.get(timeout=timeout)
    except queue.Empty:
        return "timeout"


def test_register_and_node_capacity(world):
    p = world.make()
    p.Serve(world.kubelet.socket)
    req = world.kubelet.register_requests.get(timeout=5)
    assert req.hex() == KAT["register_request_hex"]
    node = world.kube.nodes[NODE]
    assert node["status"]["capacity"]["aliyun.com/gpu-count"] == "8" == node["status"]["allocatable"]["aliyun.com/gpu-count"]
    patches = [r for r in world.kube.requests if r[0] == "PATCH"]
    assert patches[0][1] == f"/api/v1/nodes/{NODE}/status" and patches[0][3] == "application/strategic-merge-patch+json"
    ch = world.kubelet.channel("aliyungpushare.sock")
    assert world.kubelet.get_options(ch) == b"" and world.kubelet.pre_start(ch) == b""
    ch.close()
    # second plugin on the same node: capacity already right -> no PATCH (podmanager.go:80-85)
    n_before = len([r for r in world.kube.requests if r[0] == "PATCH"])
    p.Stop()
    assert not os.path.exists(p.socket)  # stale socket removed (server.go:195-201)
    world.make()
    assert len([r for r in world.kube.requests if r[0] == "PATCH"]) == n_before


@pytest.mark.parametrize("coalesce", [True, False], ids=["coalesced", "reference-stream"])
def test_list_and_watch_stream(world, coalesce):
    p = world.make(coalesce_health=coalesce)
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    it = Frames(world.kubelet.list_and_watch(ch))
    first = next_frame(it)
    assert first == wo.marshal_ListAndWatchResponse(all_devs()) and len(first) == 83608
    # application-error XIDs leave the GPU healthy (nvidia.go:134)
    device.health_inject(fakes.UUIDS[5], 8, 31)
    assert next_frame(it, 0.7) == "timeout"
    # a critical XID flips every fake device of that GPU
    device.health_inject(fakes.UUIDS[5], 8, 79)
    ids = [d[0] for d in all_devs()]
    hits = wo.xid_event_effects(ids, 8, 79, fakes.UUIDS[5])
    want = wo.list_and_watch_stream(all_devs(), hits)
    if coalesce:
        frames = [next_frame(it)]
        while wo.unmarshal_ListAndWatchResponse(frames[-1]) != wo.unmarshal_ListAndWatchResponse(want[-1]):
            frames.append(next_frame(it))  # events may straddle a wake-up; the end state is what counts
        assert frames[-1] == want[-1] and len(frames) < 20
    else:
        got = [next_frame(it) for _ in range(179)]
        assert got == want[1:]  # the reference's exact stream: one full list per fake device
    # Unhealthy is sticky and an event without UUID takes every device down (nvidia.go:138-144)
    device.health_inject("", 8, 48)
    last = None
    deadline = time.monotonic() + 10
    while time.monotonic() < deadline:
        f = next_frame(it, 2)
        if f == "timeout":
            break
        last = f
    assert all(h == wo.Unhealthy for _, h in wo.unmarshal_ListAndWatchResponse(last))
    p.Stop()
    assert next_frame(it, 5) is None  # the stream ends with the server (ListAndWatch returns nil on stop)
    ch.close()


def test_probe_events_mark_the_gpu_unhealthy(world):
    p = world.make()
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    it = Frames(world.kubelet.list_and_watch(ch))
    next_frame(it)
    device.health_inject(fakes.UUIDS[2], 0x100, 1)  # GSB_EVENT_PROBE / mismatch
    f = next_frame(it)
    devs = wo.unmarshal_ListAndWatchResponse(f)
    bad = {wo.extractRealDeviceID(i) for i, h in devs if h == wo.Unhealthy}
    assert bad == {fakes.UUIDS[2]}
    ch.close()


def test_allocate_config4_end_to_end(world):
    p = world.make()
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    minor_of = dict(zip(range(8), fakes.MINORS))
    t0 = time.time_ns()
    for i in range(64):
        req = wo.marshal_AllocateRequest([[wo.generateFakeDeviceID(fakes.UUIDS[(i * 3) % 8], j) for j in range(4)]])
        envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req))
        idx = i // 8  # the pod's ALIYUN_COM_GPU_MEM_IDX annotation; must be a /dev/nvidia MINOR on this node
        assert idx in minor_of.values()
        assert envs == [{"NVIDIA_VISIBLE_DEVICES": str(idx), "ALIYUN_COM_GPU_MEM_IDX": str(idx),
                         "ALIYUN_COM_GPU_MEM_POD": "4", "ALIYUN_COM_GPU_MEM_CONTAINER": "4",
                         "ALIYUN_COM_GPU_MEM_DEV": "179"}]
        ann = world.kube.pod(f"pod-{i:02d}")["metadata"]["annotations"]
        assert ann["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true" and int(ann["ALIYUN_COM_GPU_MEM_ASSUME_TIME"]) >= t0
    # every pod is assigned now: the 65th request gets the poison envs, gRPC status still OK
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))
    assert envs == [KAT["err_response"]["envs"]]
    pod_patches = [r for r in world.kube.requests if r[0] == "PATCH" and "/pods/" in r[1]]
    assert len(pod_patches) == 64 and pod_patches[0][3] == "application/strategic-merge-patch+json"
    assert re.fullmatch(rb'\{"metadata":\{"annotations":\{"ALIYUN_COM_GPU_MEM_ASSIGNED":"true",'
                        rb'"ALIYUN_COM_GPU_MEM_ASSUME_TIME":"\d{19}"\}\}\}', pod_patches[0][2])
    lists = [r for r in world.kube.requests if r[0] == "GET" and r[1].startswith("/api/v1/pods?")]
    assert "spec.nodeName%3Db200-0%2Cstatus.phase%3DPending" in lists[0][1]
    ch.close()


def test_allocate_failure_paths_never_raise(world):
    p = world.make()
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
    err = [KAT["err_response"]["envs"]]

    def n_patches():
        return len([r for r in world.kube.requests if r[0] == "PATCH" and "/pods/" in r[1]])
    # optimistic-lock conflict: exactly one retry, then success (allocate.go:138-144)
    world.kube.fail_next_patch(const.OptimisticLockErrorMsg, 1)
    before = n_patches()
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req))
    assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0" and n_patches() == before + 2
    # two conflicts in a row: error envs
    world.kube.fail_next_patch(const.OptimisticLockErrorMsg, 2)
    before = n_patches()
    assert wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req)) == err and n_patches() == before + 2
    # any other PATCH error: no retry
    world.kube.fail_next_patch("pods is forbidden", 1)
    before = n_patches()
    assert wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req)) == err and n_patches() == before + 1
    # LIST keeps failing (1 try + 3 retries): error envs (allocate.go:62-66)
    world.kube.fail_lists = 4
    assert wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req)) == err
    # LIST fails three times then works: the request is served
    world.kube.fail_lists = 3
    assert wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
    # bytes gogo's Unmarshal refuses: the call fails with INTERNAL as under grpc-go; nothing is LISTed or PATCHed
    before = len(world.kube.requests)
    for bad in (b"\x0a\x05\x0a", b"\x08\x01", b"\x0c"):
        with pytest.raises(grpc.RpcError) as e:
            world.kubelet.allocate(ch, bad)
        assert e.value.code() == grpc.StatusCode.INTERNAL and "error unmarshalling request" in e.value.details()
    assert len(world.kube.requests) == before
    ch.close()


def test_query_kubelet_path_and_cgpu_label(world):
    world.kube.nodes[NODE]["metadata"]["labels"]["cgpu.disable.isolation"] = "true"
    kc = NewKubeletClient(KubeletClientConfig(Address="127.0.0.1", Port=world.kube.port, BearerToken="t", Scheme="http"))
    p = world.make(queryKubelet=True, client=kc)
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))
    assert envs[0]["CGPU_DISABLE"] == "true" and envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
    assert any(r[:2] == ("GET", "/pods/") for r in world.kube.requests)
    ch.close()


def test_single_gpu_shortcut_over_grpc(world, monkeypatch):
    fakes.install(monkeypatch, n_gpus=1)
    for k in list(world.kube.pods):
        world.kube.pods[k]["status"]["phase"] = "Running"  # no pending candidates
    p = world.make()
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))
    assert envs == [{"NVIDIA_VISIBLE_DEVICES": fakes.UUIDS[0], "ALIYUN_COM_GPU_MEM_IDX": str(fakes.MINORS[0]),
                     "ALIYUN_COM_GPU_MEM_POD": "2", "ALIYUN_COM_GPU_MEM_CONTAINER": "2", "ALIYUN_COM_GPU_MEM_DEV": "179"}]
    ch.close()


def test_register_fails_without_kubelet(world):
    world.kubelet.stop()
    p = world.make()
    with pytest.raises(Exception):
        p.Serve(world.kubelet.socket)
    assert p.server is None and not os.path.exists(p.socket)  # Serve -> Stop on register failure (server.go:233-237)


# ---- §8(f) row 2: pending-pod cache + narrowed lock -------------------------------------------------

def _lists(world):
    return len([r for r in world.kube.requests if r[0] == "GET" and r[1].startswith("/api/v1/pods?")])


def test_ttl_zero_lists_on_every_call_like_the_reference(world):
    p = world.make(pod_cache_ttl=0)
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
    for i in range(10):
        assert wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == str(i // 8)
    assert _lists(world) == 10
    ch.close()


def test_cache_skips_the_list_but_refreshes_on_a_miss(world):
    p = world.make(pod_cache_ttl=60)
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
    for i in range(64):
        assert wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == str(i // 8)
    assert _lists(world) == 1  # one LIST served 64 requests
    # a pod the scheduler bound AFTER the cache was filled: first look misses, the refresh finds it
    new = make_pod(99, NODE, gpu_mem=2, idx=5, assume_time=1_800_000_000_000_000_000)
    with world.kube.lock:
        world.kube.pods[("default", "pod-99")] = new
        world.kube.order.append(("default", "pod-99"))
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))
    assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "5" and _lists(world) == 2
    assert world.kube.pod("pod-99")["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true"
    # nothing left: the miss costs one more LIST, the answer is the reference's poison envs
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, req))
    assert envs == [KAT["err_response"]["envs"]] and _lists(world) == 3
    ch.close()


def test_concurrent_allocates_never_share_a_pod(world):
    p = world.make(pod_cache_ttl=60, max_workers=32)
    p.Serve(world.kubelet.socket)
    results, errors = [], []

    def client():
        try:
            ch = world.kubelet.channel("aliyungpushare.sock")
            for _ in range(4):
                results.append(wo.unmarshal_AllocateResponse(
                    world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))[0])
            ch.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)
    ts = [threading.Thread(target=client) for _ in range(16)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert not errors and len(results) == 64
    assert sorted(int(r["ALIYUN_COM_GPU_MEM_IDX"]) for r in results) == sorted(i // 8 for i in range(64))
    patched = [r[1] for r in world.kube.requests if r[0] == "PATCH" and "/pods/" in r[1]]
    assert len(patched) == 64 and len(set(patched)) == 64  # every pod claimed exactly once


@pytest.mark.parametrize("ttl", [0, 60])
def test_a_list_taken_before_a_patch_lands_cannot_hand_out_the_pod_twice(world, ttl):
    """The lock is not held across the PATCH, so a later LIST can still show a pod as unassigned while this
    plugin's PATCH for it is in flight; claims are kept by UID across every table rebuild."""
    world.kube.patch_delay = 0.25
    p = world.make(pod_cache_ttl=ttl, max_workers=32)
    p.Serve(world.kubelet.socket)
    results = []

    def one():
        ch = world.kubelet.channel("aliyungpushare.sock")
        results.append(wo.unmarshal_AllocateResponse(
            world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))[0])
        ch.close()
    ts = [threading.Thread(target=one) for _ in range(12)]
    for t in ts:
        t.start()
        threading.Event().wait(0.03)  # staggered: each request decides while earlier PATCHes are in flight
    [t.join(30) for t in ts]
    assert len(results) == 12 and all(r["ALIYUN_COM_GPU_MEM_IDX"] != "-1" for r in results)
    patched = [r[1] for r in world.kube.requests if r[0] == "PATCH" and "/pods/" in r[1]]
    assert len(patched) == 12 and len(set(patched)) == 12


@pytest.mark.parametrize("ttl", [0, 0.05, 60])
def test_everything_at_once_for_a_few_seconds(world, ttl):
    """Soak of the Python front end: concurrent Allocates, pod churn on the apiserver, health events and
    ListAndWatch streams coming and going. No applied PATCH may repeat, the plugin keeps answering, and the final
    stream shows exactly the marked GPU."""
    p = world.make(pod_cache_ttl=ttl, max_workers=32)
    p.Serve(world.kubelet.socket)
    stop, errors, results = threading.Event(), [], []
    ids = iter(range(1000, 100000))
    pause = threading.Event().wait  # time.sleep is stubbed out by the fixture

    def guard(fn):
        def run():
            try:
                while not stop.is_set():
                    fn()
            except Exception as e:  # noqa: BLE001
                if not stop.is_set():
                    errors.append(repr(e))
        return run

    def allocator():
        ch = world.kubelet.channel("aliyungpushare.sock")
        r = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))[0]
        results.append(r["ALIYUN_COM_GPU_MEM_IDX"])
        ch.close()

    def churn():
        i = next(ids)
        world.kube.add_pod(make_pod(i, NODE, gpu_mem=2, idx=i % 8, assume_time=1_800_000_000_000_000_000 + i))
        if i % 3 == 0:
            try:
                world.kube.delete_pod(f"pod-{i - 2}")
            except KeyError:
                pass
        pause(0.01)

    def watcher():
        ch = world.kubelet.channel("aliyungpushare.sock")
        call = world.kubelet.list_and_watch(ch)
        next(iter(call))
        pause(0.05)
        call.cancel()
        ch.close()

    def health():
        device.health_inject(fakes.UUIDS[3], 8, 31)  # benign XID
        device.health_inject(fakes.UUIDS[6], 0x100, 1)
        pause(0.05)

    ts = [threading.Thread(target=guard(f)) for f in (churn, watcher, health, allocator, allocator, allocator)]
    [t.start() for t in ts]
    pause(3.0)
    stop.set()
    [t.join(30) for t in ts]
    assert not errors, errors[:3]
    applied = list(world.kube.patched_ok)
    assert len(results) > 30 and len(applied) == len(set(applied))
    assert len([r for r in results if r != "-1"]) <= len(applied) <= len([r for r in results if r != "-1"]) + 3
    ch = world.kubelet.channel("aliyungpushare.sock")
    devs = wo.unmarshal_ListAndWatchResponse(next(iter(world.kubelet.list_and_watch(ch))))
    assert {wo.extractRealDeviceID(i) for i, h in devs if h == wo.Unhealthy} == {fakes.UUIDS[6]}
    ch.close()


def test_optional_recovery_is_off_by_default_and_flag_gated(world):
    """server.go:180 FIXME: the reference never leaves Unhealthy. Default: a RECOVERED probe event changes
    nothing; with health_recovery_cycles > 0 the GPU's fake devices flip back — but only after a PROBE
    fault, never after an XID."""
    p = world.make()
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    it = Frames(world.kubelet.list_and_watch(ch))
    next_frame(it)
    device.health_inject(fakes.UUIDS[1], 0x100, 1)
    bad = next_frame(it)
    while True:
        f = next_frame(it, 1)
        if f == "timeout":
            break
        bad = f
    assert {wo.extractRealDeviceID(i) for i, h in wo.unmarshal_ListAndWatchResponse(bad) if h == wo.Unhealthy} == {fakes.UUIDS[1]}
    device.health_inject(fakes.UUIDS[1], 0x100, 3)  # GSB_PROBE_RECOVERED
    assert next_frame(it, 0.8) == "timeout"          # sticky, like the reference
    ch.close()
    p.Stop()
    p2 = world.make(health_recovery_cycles=3)
    p2.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    it = Frames(world.kubelet.list_and_watch(ch))
    first = next_frame(it)
    def settle(want):  # per-device events may straddle a wake-up of the stream: the end state is what counts
        for _ in range(200):
            f = next_frame(it)
            if f == want:
                return True
            assert f != "timeout"
        return False
    device.health_inject(fakes.UUIDS[1], 0x100, 1)
    assert settle(bad)
    device.health_inject(fakes.UUIDS[1], 0x100, 3)
    assert settle(first)  # all Healthy again, byte-identical to the initial list
    ch.close()


# ---- round-2 advisor findings ---------------------------------------------------------------------------

def test_health_event_raised_before_any_stream_is_in_the_first_frame(world):
    """The reference's unbuffered channel blocks the producer until a stream takes the event, so nothing is lost;
    here the producer writes the node state itself, so a fault raised before the kubelet opens ListAndWatch (a
    start-up walk fault, a NOT_SUPPORTED registration, a reconnect window) is in that stream's first frame."""
    p = world.make()
    p.Serve(world.kubelet.socket)
    device.health_inject(fakes.UUIDS[3], 8, 79)
    deadline = time.monotonic() + 5
    while p.devs[3 * 179 + 178].Health != const.Unhealthy and time.monotonic() < deadline:  # the LAST fake device of GPU 3
        threading.Event().wait(0.02)
    ch = world.kubelet.channel("aliyungpushare.sock")
    it = Frames(world.kubelet.list_and_watch(ch))
    assert wo.unmarshal_ListAndWatchResponse(next_frame(it)) == all_devs(unhealthy={3})
    assert next_frame(it, 0.6) == "timeout"  # and it is not re-sent as news
    # a second stream (kubelet reconnect) starts from the same state; a later event reaches both
    it2 = Frames(world.kubelet.list_and_watch(ch))
    assert wo.unmarshal_ListAndWatchResponse(next_frame(it2)) == all_devs(unhealthy={3})
    device.health_inject(fakes.UUIDS[6], 0x100, 1)
    for s in (it, it2):  # this front end marks fake devices one at a time: coalescing may take more than one frame
        last, f = None, next_frame(s)
        while f != "timeout":
            last, f = f, next_frame(s, 1.0)
        assert wo.unmarshal_ListAndWatchResponse(last) == all_devs(unhealthy={3, 6})
    ch.close()


def test_single_gpu_node_does_not_answer_a_stale_cache_with_the_shortcut(world, monkeypatch):
    """One GPU + a cached pod table that pre-dates the pod being started: the reference LISTs on every call, finds
    the pod, answers with its annotated index and PATCHes it. A cached table must not let the single-GPU shortcut
    (allocate.go:151-177) answer instead — the pod would stay unassigned and be mis-matched later."""
    fakes.install(monkeypatch, n_gpus=1)
    for k in list(world.kube.pods):
        world.kube.pods[k]["status"]["phase"] = "Running"
    p = world.make(pod_cache_ttl=60)
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    # fills the cache; nothing pending -> the shortcut is the right answer here
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c"]])))
    assert envs[0]["NVIDIA_VISIBLE_DEVICES"] == fakes.UUIDS[0] and _lists(world) == 1  # decided on a fresh LIST
    new = make_pod(99, NODE, gpu_mem=2, idx=fakes.MINORS[0], assume_time=1_800_000_000_000_000_000)
    with world.kube.lock:
        world.kube.pods[("default", "pod-99")] = new
        world.kube.order.append(("default", "pod-99"))
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))
    assert envs[0]["NVIDIA_VISIBLE_DEVICES"] == str(fakes.MINORS[0]) == envs[0]["ALIYUN_COM_GPU_MEM_IDX"]
    assert world.kube.pod("pod-99")["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true"
    ch.close()


def test_inventory_event_marks_only_that_gpu_unhealthy(world):
    """GSB_EVENT_INVENTORY from the library's off-path NVML refresh: that GPU's fake devices go Unhealthy, the others
    stay; unknown event types are ignored (nvidia.go:127-129)."""
    p = world.make()
    p.Serve(world.kubelet.socket)
    ch = world.kubelet.channel("aliyungpushare.sock")
    it = Frames(world.kubelet.list_and_watch(ch))
    assert next_frame(it) == wo.marshal_ListAndWatchResponse(all_devs())
    device.health_inject(fakes.UUIDS[4], 0x200, 2)
    last, f = None, next_frame(it)
    while f != "timeout":
        last, f = f, next_frame(it, 1.0)
    assert wo.unmarshal_ListAndWatchResponse(last) == all_devs(unhealthy={4})
    device.health_inject(fakes.UUIDS[4], 0x300, 1)
    assert next_frame(it, 0.7) == "timeout"
    ch.close()
