"""gsbd — the native (C++) daemon — end to end on CPU with a synthetic 8-GPU inventory: real HTTP/2
from grpcio (fake kubelet: Registration server + device-manager client), the stateful mock apiserver,
signals and inotify. Same expectations as tests/test_server.py has of the Python front end; wire
bytes are checked against the oracle's restatement of the reference."""
import json
import os
import re
import signal
import subprocess
import threading
import time

import grpc
import pytest

from gpushare_device_plugin_b200.testing.fake_kubelet import FakeKubelet
from gpushare_device_plugin_b200.testing.mock_kube import MockKube, config4_pods, make_node, make_pod
from oracle import wire_oracle as wo

from . import fakes
from .test_server import Frames, next_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GSBD = os.environ.get("GSBD_BINARY") or os.path.join(ROOT, "gpushare_device_plugin_b200", "gsbd")  # GSBD_BINARY: a sanitizer build
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "wire_kat.json")))
NODE = "b200-0"

pytestmark = pytest.mark.skipif(not os.access(GSBD, os.X_OK), reason="gsbd not built")


class Daemon:
    def __init__(self, tmp_path, kube, *extra, n_gpus=8, wait_register=True, kubelet=None):
        self.dir = tmp_path
        self.kubelet = kubelet or FakeKubelet(str(tmp_path))
        env = dict(os.environ, NODE_NAME=NODE, GPUSHARE_PLUGIN_DIR=str(tmp_path) + "/", GPUSHARE_DUMP_DIR=str(tmp_path),
                   GPUSHARE_RETRY_SLEEP_MS="1", GSBD_ALLOW_FAKE_INVENTORY="1")
        env.pop("KUBECONFIG", None)
        self.log = open(tmp_path / "gsbd.log", "w")
        self.proc = subprocess.Popen([GSBD, "-logtostderr", "--v=5", "--memory-unit=GiB", "--health-check",
                                      "--fake-inventory", str(n_gpus), "--kube-api-url", kube.url, *extra],
                                     env=env, stderr=self.log, stdout=self.log)
        self.register_request = self.kubelet.register_requests.get(timeout=20) if wait_register else None

    def channel(self):
        return self.kubelet.channel("aliyungpushare.sock")

    def inject(self, ch, uuid, etype, edata):
        ch.unary_unary("/gsbd.Test/InjectEvent")(f"{uuid or '-'} {etype} {edata}".encode(), timeout=5)

    def close(self):
        if self.proc.poll() is None:
            self.proc.send_signal(signal.SIGTERM)
            try:
                self.proc.wait(timeout=10)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.log.close()
        self.kubelet.stop()


@pytest.fixture
def world(tmp_path):
    kube = MockKube(make_node(NODE), config4_pods(NODE))
    daemons = []

    def start(*extra, **kw):
        d = Daemon(tmp_path, kube, *extra, **kw)
        daemons.append(d)
        return d
    yield type("W", (), {"kube": kube, "start": staticmethod(start), "dir": tmp_path})
    for d in daemons:
        d.close()
    kube.close()
    print(open(tmp_path / "gsbd.log").read()[-2500:])


def all_devs(unhealthy=()):
    return [[wo.generateFakeDeviceID(u, j), wo.Unhealthy if g in unhealthy else wo.Healthy]
            for g, u in enumerate(fakes.UUIDS) for j in range(179)]


def test_register_node_capacity_and_trivial_rpcs(world):
    d = world.start()
    assert d.register_request.hex() == KAT["register_request_hex"]
    node = world.kube.nodes[NODE]
    assert node["status"]["capacity"]["aliyun.com/gpu-count"] == "8" == node["status"]["allocatable"]["aliyun.com/gpu-count"]
    patch = [r for r in world.kube.requests if r[0] == "PATCH"][0]
    assert patch[1] == f"/api/v1/nodes/{NODE}/status" and patch[3] == "application/strategic-merge-patch+json"
    assert json.loads(patch[2]) == {"status": {"allocatable": {"aliyun.com/gpu-count": "8"},
                                               "capacity": {"aliyun.com/gpu-count": "8"}}}
    ch = d.channel()
    assert d.kubelet.get_options(ch) == b"" and d.kubelet.pre_start(ch) == b""
    with pytest.raises(grpc.RpcError) as e:
        ch.unary_unary("/v1beta1.DevicePlugin/Nope")(b"", timeout=5)
    assert e.value.code() == grpc.StatusCode.UNIMPLEMENTED
    ch.close()


@pytest.mark.parametrize("coalesce", [True, False], ids=["coalesced", "reference-stream"])
def test_list_and_watch_stream(world, coalesce):
    d = world.start("--coalesce-health=" + ("true" if coalesce else "false"))
    ch = d.channel()
    it = Frames(d.kubelet.list_and_watch(ch))
    first = next_frame(it)
    assert first == wo.marshal_ListAndWatchResponse(all_devs()) and len(first) == 83608
    d.inject(ch, fakes.UUIDS[5], 8, 31)  # application-error XID: stays healthy (nvidia.go:134)
    assert next_frame(it, 0.7) == "timeout"
    d.inject(ch, fakes.UUIDS[5], 8, 79)
    ids = [x[0] for x in all_devs()]
    want = wo.list_and_watch_stream(all_devs(), wo.xid_event_effects(ids, 8, 79, fakes.UUIDS[5]))
    if coalesce:
        assert next_frame(it) == want[-1]  # one resend carrying all 179 flips
    else:
        assert [next_frame(it) for _ in range(179)] == want[1:]  # the reference's exact stream
    d.inject(ch, fakes.UUIDS[2], 0x100, 1)  # active-probe verdict
    last = next_frame(it)
    while True:
        f = next_frame(it, 1)
        if f == "timeout":
            break
        last = f
    assert wo.unmarshal_ListAndWatchResponse(last) == all_devs(unhealthy={2, 5})
    d.inject(ch, "", 8, 48)  # no UUID: every device (nvidia.go:138-144)
    last = next_frame(it, 10)
    while True:
        f = next_frame(it, 1)
        if f == "timeout":
            break
        last = f
    assert all(h == wo.Unhealthy for _, h in wo.unmarshal_ListAndWatchResponse(last))
    ch.close()


def test_allocate_config4_end_to_end(world):
    d = world.start()
    ch = d.channel()
    t0 = time.time_ns()
    for i in range(64):
        req = wo.marshal_AllocateRequest([[wo.generateFakeDeviceID(fakes.UUIDS[(i * 3) % 8], j) for j in range(4)]])
        raw = d.kubelet.allocate(ch, req)
        envs = wo.unmarshal_AllocateResponse(raw)
        assert envs == [{"NVIDIA_VISIBLE_DEVICES": str(i // 8), "ALIYUN_COM_GPU_MEM_IDX": str(i // 8),
                         "ALIYUN_COM_GPU_MEM_POD": "4", "ALIYUN_COM_GPU_MEM_CONTAINER": "4",
                         "ALIYUN_COM_GPU_MEM_DEV": "179"}]
        assert raw == wo.marshal_AllocateResponse(envs)
        ann = world.kube.pod(f"pod-{i:02d}")["metadata"]["annotations"]
        assert ann["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true" and int(ann["ALIYUN_COM_GPU_MEM_ASSUME_TIME"]) >= t0
    envs = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))
    assert envs == [KAT["err_response"]["envs"]]
    pod_patches = [r for r in world.kube.requests if r[0] == "PATCH" and "/pods/" in r[1]]
    assert len(pod_patches) == 64 and pod_patches[0][3] == "application/strategic-merge-patch+json"
    assert re.fullmatch(rb'\{"metadata":\{"annotations":\{"ALIYUN_COM_GPU_MEM_ASSIGNED":"true",'
                        rb'"ALIYUN_COM_GPU_MEM_ASSUME_TIME":"\d{19}"\}\}\}', pod_patches[0][2])
    lists = [r for r in world.kube.requests if r[0] == "GET" and r[1].startswith("/api/v1/pods?")]
    assert "spec.nodeName%3Db200-0%2Cstatus.phase%3DPending" in lists[0][1] and len(lists) <= 4  # pod cache
    ch.close()


def test_allocate_failure_paths_and_cache(world):
    d = world.start("--pod-cache-ttl", "60", "--pod-informer=false")  # the TTL cache alone (LIST failures are injected)
    ch = d.channel()
    req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
    err = [KAT["err_response"]["envs"]]

    def n(kind, sub):
        return len([r for r in world.kube.requests if r[0] == kind and sub in r[1]])
    world.kube.fail_next_patch("the object has been modified; please apply your changes to the latest version and try again", 1)
    before = n("PATCH", "/pods/")
    assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
    assert n("PATCH", "/pods/") == before + 2  # exactly one retry (allocate.go:138-144)
    world.kube.fail_next_patch("pods is forbidden", 1)
    before = n("PATCH", "/pods/")
    assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req)) == err and n("PATCH", "/pods/") == before + 1
    world.kube.fail_lists = 4  # cache was dropped by the failed PATCH: 1 try + 3 retries all fail
    assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req)) == err
    world.kube.fail_lists = 3
    assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
    lists = n("GET", "/api/v1/pods?")
    for _ in range(10):
        d.kubelet.allocate(ch, req)
    assert n("GET", "/api/v1/pods?") == lists  # served from the cache
    new = make_pod(99, NODE, gpu_mem=2, idx=5, assume_time=1_800_000_000_000_000_000)
    with world.kube.lock:
        world.kube.pods[("default", "pod-99")] = new
        world.kube.order.append(("default", "pod-99"))
    envs = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))
    assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "5" and n("GET", "/api/v1/pods?") == lists + 1  # refresh on a miss
    # bytes gogo's Unmarshal refuses: the call fails with INTERNAL as under grpc-go, and nothing was LISTed or PATCHed
    before = len(world.kube.requests)
    for bad in (b"\x0a\x05\x0a", b"\x08\x01", b"\x0c"):
        with pytest.raises(grpc.RpcError) as e:
            d.kubelet.allocate(ch, bad)
        assert e.value.code() == grpc.StatusCode.INTERNAL and "error unmarshalling request" in e.value.details()
    assert len(world.kube.requests) == before
    # unknown fields are skipped like any proto3 reader does: still answered
    ok = wo.marshal_AllocateRequest([["a", "b"]]) + b"\x10\x07"
    assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, ok))[0]["ALIYUN_COM_GPU_MEM_CONTAINER"] == "2"
    ch.close()


def test_concurrent_allocates_never_share_a_pod(world):
    d = world.start("--pod-cache-ttl", "60")
    results, errors = [], []

    def client():
        try:
            ch = d.channel()
            for _ in range(4):
                results.append(wo.unmarshal_AllocateResponse(
                    d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))[0])
            ch.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)
    ts = [threading.Thread(target=client) for _ in range(16)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert not errors and len(results) == 64
    assert sorted(int(r["ALIYUN_COM_GPU_MEM_IDX"]) for r in results) == sorted(i // 8 for i in range(64))
    patched = [r[1] for r in world.kube.requests if r[0] == "PATCH" and "/pods/" in r[1]]
    assert len(patched) == 64 and len(set(patched)) == 64


def test_query_kubelet_cgpu_label_and_single_gpu(world):
    world.kube.nodes[NODE]["metadata"]["labels"]["cgpu.disable.isolation"] = "true"
    d = world.start("--query-kubelet", "--kubelet-address", "127.0.0.1", "--kubelet-port", str(world.kube.port),
                    "--kubelet-scheme", "http", "--token", "t")
    ch = d.channel()
    envs = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))
    assert envs[0]["CGPU_DISABLE"] == "true" and envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
    assert any(r[:2] == ("GET", "/pods/") for r in world.kube.requests)
    ch.close()
    d.close()
    for k in list(world.kube.pods):
        world.kube.pods[k]["status"]["phase"] = "Running"
    d1 = world.start(n_gpus=1)
    ch = d1.channel()
    envs = wo.unmarshal_AllocateResponse(d1.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))
    assert envs == [{"NVIDIA_VISIBLE_DEVICES": fakes.UUIDS[0], "ALIYUN_COM_GPU_MEM_IDX": str(fakes.MINORS[0]),
                     "ALIYUN_COM_GPU_MEM_POD": "2", "ALIYUN_COM_GPU_MEM_CONTAINER": "2", "ALIYUN_COM_GPU_MEM_DEV": "179",
                     "CGPU_DISABLE": "true"}]
    ch.close()


def test_lifecycle_restart_dump_and_exit_codes(world, tmp_path):
    d = world.start()
    ch = d.channel()
    assert len(wo.unmarshal_ListAndWatchResponse(next(iter(d.kubelet.list_and_watch(ch))))) == 1432
    ch.close()
    d.proc.send_signal(signal.SIGQUIT)  # dump, keep running (gpumanager.go:97-101)
    deadline = time.time() + 5
    while time.time() < deadline and not [f for f in os.listdir(tmp_path) if f.startswith("go_")]:
        time.sleep(0.05)
    dumps = [f for f in os.listdir(tmp_path) if f.startswith("go_")]
    assert dumps and "gpus 8 slices 179" in open(tmp_path / dumps[0]).read() and d.proc.poll() is None
    d.kubelet.stop()  # kubelet restart: socket re-created -> rebuild + re-register (gpumanager.go:83-87)
    time.sleep(0.2)
    d.kubelet.start()
    assert d.kubelet.register_requests.get(timeout=20) == d.register_request
    d.proc.send_signal(signal.SIGHUP)  # same via SIGHUP (:94-96)
    assert d.kubelet.register_requests.get(timeout=20) == d.register_request
    d.proc.send_signal(signal.SIGTERM)
    assert d.proc.wait(timeout=10) == 0 and not os.path.exists(tmp_path / "aliyungpushare.sock")
    # no kubelet to register with: exit status 2 (gpumanager.go:76)
    d.kubelet.stop()
    d2 = Daemon(tmp_path, world.kube, wait_register=False, kubelet=d.kubelet)
    assert d2.proc.wait(timeout=20) == 2
    d2.log.close()
    # NODE_NAME missing: fatal at kubeInit
    env = dict(os.environ, GSBD_ALLOW_FAKE_INVENTORY="1")
    env.pop("NODE_NAME", None)
    p = subprocess.run([GSBD, "--fake-inventory", "1", "--kube-api-url", world.kube.url], env=env, capture_output=True, text=True)
    assert p.returncode != 0 and "Please set env NODE_NAME" in p.stderr
    env = dict(os.environ, NODE_NAME=NODE)
    env.pop("GSBD_ALLOW_FAKE_INVENTORY", None)  # the synthetic-inventory hook cannot be switched on by the flag alone
    p = subprocess.run([GSBD, "--fake-inventory", "1", "--kube-api-url", world.kube.url], env=env, capture_output=True, text=True)
    assert p.returncode == 2 and "test hook" in p.stderr
    p = subprocess.run([GSBD, "--no-such-flag"], capture_output=True, text=True)
    assert p.returncode == 2 and "flag provided but not defined" in p.stderr


def test_chunked_lists_and_awkward_json(tmp_path):
    """The real apiserver answers LISTs chunked; pods carry escapes, non-ASCII text, numeric and
    suffixed quantities, nulls and nested objects. The native JSON/HTTP code must read them exactly like
    the reference's decoder: same pod picked, same envs."""
    pods = config4_pods(NODE, 8)
    pods[0]["metadata"]["annotations"]["note"] = 'quote " backslash \\ tab \t newline \n unicode é 漢字 \U0001F600'
    pods[0]["metadata"]["labels"] = {"a": None, "deep": {"x": [1, 2.5, -3e2, True, False, None, {"y": []}]}}
    pods[0]["spec"]["containers"][0]["resources"]["limits"]["aliyun.com/gpu-mem"] = 3          # a JSON number
    pods[1]["spec"]["containers"][0]["resources"]["limits"]["aliyun.com/gpu-mem"] = "2k"       # 2000
    pods[2]["spec"]["containers"][0]["resources"]["limits"]["aliyun.com/gpu-mem"] = "1500m"    # rounds up to 2
    pods[2]["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_IDX"] = "6"
    pods[3]["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSUME_TIME"] = "12x"               # unparsable -> 0 = oldest
    pods[3]["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_IDX"] = "7"
    kube = MockKube(make_node(NODE), pods, chunked_lists=True)
    d = Daemon(tmp_path, kube, "--pod-cache-ttl", "0")
    try:
        ch = d.channel()
        minors = {u: m for u, m in zip(fakes.UUIDS, fakes.MINORS)}
        for n_ids in (3, 2, 4, 2000):
            req = [["x"] * n_ids]
            want, want_pod = wo.Allocate(req, [kube.pod(f"pod-{i:02d}") for i in range(8)], NODE, minors, 179)
            got = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest(req)))
            assert got == want, (n_ids, got, want)
            if want_pod is not None:
                assert kube.pod(want_pod["metadata"]["name"])["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true"
        ch.close()
    finally:
        d.close()
        kube.close()


def test_native_optional_recovery(world):
    d = world.start()
    ch = d.channel()
    it = Frames(d.kubelet.list_and_watch(ch))
    first = next_frame(it)
    d.inject(ch, fakes.UUIDS[1], 0x100, 1)
    bad = next_frame(it)
    d.inject(ch, fakes.UUIDS[1], 0x100, 3)
    assert next_frame(it, 0.8) == "timeout"  # default: sticky like the reference
    ch.close()
    d.close()
    d2 = world.start("--health-recovery-cycles", "3")
    ch = d2.channel()
    it = Frames(d2.kubelet.list_and_watch(ch))
    assert next_frame(it) == first
    d2.inject(ch, fakes.UUIDS[1], 0x100, 1)
    assert next_frame(it) == bad
    d2.inject(ch, fakes.UUIDS[1], 0x100, 3)
    assert next_frame(it) == first
    ch.close()


def test_native_load_generator_drives_config4(world):
    """csrc/daemon/alloc_load.cc (persistent HTTP/2 client) against gsbd: 64 requests, 64 distinct pods claimed."""
    lg = os.path.join(ROOT, "gpushare_device_plugin_b200", "gsb_alloc_load")
    if not os.access(lg, os.X_OK):
        pytest.skip("gsb_alloc_load not built")
    d = world.start("--pod-cache-ttl", "60")
    out = subprocess.run([lg, str(world.dir / "aliyungpushare.sock"), "4", "64", ",".join(fakes.UUIDS)],
                         capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout)
    assert (r["requests"], r["error_responses"], r["rpc_failures"], r["client"]) == (64, 0, 0, "native")
    assert 0 < r["p50_us"] < r["p99_us"]
    patched = [x[1] for x in world.kube.requests if x[0] == "PATCH" and "/pods/" in x[1]]
    assert len(patched) == 64 and len(set(patched)) == 64
    out = subprocess.run([lg, str(world.dir / "aliyungpushare.sock"), "1", "3", ",".join(fakes.UUIDS)],
                         capture_output=True, text=True, timeout=60)
    assert json.loads(out.stdout)["error_responses"] == 3  # nothing left to assign: poison envs, still grpc OK


def _count(world, kind, sub):
    return len([r for r in world.kube.requests if r[0] == kind and sub in r[1]])


def test_informer_keeps_the_table_current_without_lists(world):
    d = world.start("--pod-cache-ttl", "0")  # no TTL cache at all: only the watch stream can avoid a LIST per call
    ch = d.channel()
    req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
    deadline = time.time() + 5
    while world.kube.watches_served == 0 and time.time() < deadline:
        time.sleep(0.05)
    assert world.kube.watches_served >= 1
    time.sleep(0.2)
    lists0 = _count(world, "GET", "/api/v1/pods?fieldSelector")
    for i in range(64):
        assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == str(i // 8)
    assert _count(world, "GET", "/api/v1/pods?fieldSelector") == lists0  # not one LIST for 64 requests
    # a pod bound after start-up arrives as an ADDED event: found without a LIST
    world.kube.add_pod(make_pod(99, NODE, gpu_mem=2, idx=5, assume_time=1_800_000_000_000_000_000))
    time.sleep(0.3)
    envs = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))
    assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "5" and _count(world, "GET", "/api/v1/pods?fieldSelector") == lists0
    # a deleted pod disappears from the table; the miss is confirmed by one authoritative LIST
    world.kube.add_pod(make_pod(98, NODE, gpu_mem=3, idx=6, assume_time=1_800_000_000_000_000_001))
    time.sleep(0.2)
    world.kube.delete_pod("pod-98")
    time.sleep(0.3)
    envs = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c"]])))
    assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "-1" and _count(world, "GET", "/api/v1/pods?fieldSelector") == lists0 + 1
    # a failed PATCH forces a resync (fresh LIST + new watch); service continues
    world.kube.add_pod(make_pod(97, NODE, gpu_mem=5, idx=2, assume_time=1_800_000_000_000_000_002))
    time.sleep(0.3)
    world.kube.fail_next_patch("pods is forbidden", 1)
    w0 = world.kube.watches_served
    five = wo.marshal_AllocateRequest([["a"] * 5])
    assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, five))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "-1"
    deadline = time.time() + 5
    while world.kube.watches_served == w0 and time.time() < deadline:
        time.sleep(0.05)
    assert world.kube.watches_served > w0
    assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, five))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "2"
    # the apiserver expires the stream (ERROR event, 410 Gone): the informer starts over from a LIST
    w0 = world.kube.watches_served
    world.kube.expire_watches()
    deadline = time.time() + 5
    while world.kube.watches_served == w0 and time.time() < deadline:
        time.sleep(0.05)
    assert world.kube.watches_served > w0
    world.kube.add_pod(make_pod(96, NODE, gpu_mem=7, idx=3, assume_time=1_800_000_000_000_000_003))
    time.sleep(0.3)
    seven = wo.marshal_AllocateRequest([["a"] * 7])
    assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, seven))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "3"
    ch.close()


def test_informer_falls_back_to_the_ttl_cache_when_watch_is_refused(world):
    world.kube.enable_watch = False
    d = world.start("--pod-cache-ttl", "60")
    ch = d.channel()
    req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
    for i in range(16):
        assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == str(i // 8)
    ch.close()
    assert world.kube.watches_served == 0


def test_reference_exact_mode_lists_on_every_call(world):
    d = world.start("--pod-informer=false", "--pod-cache-ttl", "0")
    ch = d.channel()
    req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
    for i in range(10):
        assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == str(i // 8)
    assert _count(world, "GET", "/api/v1/pods?fieldSelector") == 10 and world.kube.watches_served == 0
    ch.close()


def test_informer_table_survives_churn(world):
    """Deletes are tombstones that get compacted away; a uid that comes back is a fresh row; order-sensitive
    decisions (oldest assume time first) stay right through all of it."""
    d = world.start("--pod-cache-ttl", "0")
    ch = d.channel()
    deadline = time.time() + 5
    while world.kube.watches_served == 0 and time.time() < deadline:
        time.sleep(0.05)
    base = 1_600_000_000_000_000_000  # older than every config-4 pod
    for i in range(100, 200):  # 100 extra pods asking for 3 GiB on GPU 7
        world.kube.add_pod(make_pod(i, NODE, gpu_mem=3, idx=7, assume_time=base + i))
    for i in range(100, 190):  # 90 of them go away again: enough tombstones to trigger a compaction
        world.kube.delete_pod(f"pod-{i}")
    world.kube.add_pod(make_pod(150, NODE, gpu_mem=3, idx=6, assume_time=base))  # same uid as a deleted pod, now oldest
    time.sleep(0.5)
    lists0 = _count(world, "GET", "/api/v1/pods?fieldSelector")
    three = wo.marshal_AllocateRequest([["a", "b", "c"]])
    got = [wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, three))[0]["ALIYUN_COM_GPU_MEM_IDX"] for _ in range(11)]
    assert got == ["6"] + ["7"] * 10  # the returned pod first, then pod-190..199 by age
    assert _count(world, "GET", "/api/v1/pods?fieldSelector") == lists0
    patched = [r[1].rsplit("/", 1)[1] for r in world.kube.requests if r[0] == "PATCH" and "/pods/" in r[1]]
    assert patched == ["pod-150"] + [f"pod-{i}" for i in range(190, 200)]
    # the original 64 pods are untouched by the churn
    four = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
    for i in range(64):
        assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, four))[0]["ALIYUN_COM_GPU_MEM_IDX"] == str(i // 8)
    ch.close()


@pytest.mark.parametrize("mode", [("--pod-informer=false", "--pod-cache-ttl", "0"), ("--pod-cache-ttl", "0"),
                                  ("--pod-informer=false", "--pod-cache-ttl", "60")])
def test_a_list_taken_before_a_patch_lands_cannot_hand_out_the_pod_twice(world, mode):
    """The lock is not held across the PATCH, so a LIST (per call, on TTL expiry, or the informer's) can show a pod
    as still unassigned while this daemon's PATCH for it is in flight. Claims survive every table rebuild."""
    world.kube.patch_delay = 0.25
    d = world.start(*mode)
    results = []

    def one():
        ch = d.channel()
        results.append(wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))[0])
        ch.close()
    ts = [threading.Thread(target=one) for _ in range(12)]
    for t in ts:
        t.start()
        time.sleep(0.03)  # staggered: each request decides while the earlier PATCHes are still in flight
    [t.join(30) for t in ts]
    assert len(results) == 12 and all(r["ALIYUN_COM_GPU_MEM_IDX"] != "-1" for r in results)
    patched = [r[1] for r in world.kube.requests if r[0] == "PATCH" and "/pods/" in r[1]]
    assert len(patched) == 12 and len(set(patched)) == 12


def test_flag_errors_follow_go_flag_package():
    """cmd/nvidia/main.go uses Go's flag package: -h prints the usage and exits 0, an undefined flag or a missing
    value prints the message plus the usage and exits 2; the ten reference flags keep their names and help texts."""
    p = subprocess.run([GSBD, "-h"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stderr.startswith("Usage of ")
    for line in ('-memory-unit string\n    \tSet memoryUnit of the GPU Memroy, support \'GiB\' and \'MiB\' (default "GiB")',
                 "-query-kubelet\n    \tQuery pending pods from kubelet instead of kube-apiserver",
                 "-kubelet-port uint\n    \tKubelet listened Port (default 10250)",
                 "-timeout int\n    \tKubelet client http timeout duration (default 10)",
                 "-mps\n    \tEnable or Disable MPS", "-health-check\n    \tEnable or disable Health check"):
        assert line in p.stderr
    p = subprocess.run([GSBD, "--no-such-flag"], capture_output=True, text=True)
    assert p.returncode == 2 and p.stderr.startswith("flag provided but not defined: -no-such-flag\nUsage of ")
    p = subprocess.run([GSBD, "-token"], capture_output=True, text=True)
    assert p.returncode == 2 and p.stderr.startswith("flag needs an argument: -token\nUsage of ")


@pytest.mark.parametrize("mode", [(), ("--pod-informer=false",), ("--pod-informer=false", "--pod-cache-ttl", "0")],
                         ids=["informer", "ttl-cache", "list-per-call"])
def test_everything_at_once_for_a_few_seconds(world, mode):
    """Soak: Allocates from several kubelet connections, pod churn on the apiserver, health events, ListAndWatch
    streams opening and closing, a kubelet restart in the middle. Afterwards: the daemon is alive and exits 0, no pod
    was handed out twice, no request was refused while a matching pod existed at the end, and a fresh stream shows
    exactly the GPUs that were marked. (tools/sanitize.sh runs this under TSan and ASan.)"""
    d = world.start(*mode)
    stop = threading.Event()
    errors, results = [], []
    next_id = [1000]
    lock = threading.Lock()

    def guard(fn):
        def run():
            try:
                while not stop.is_set():
                    fn()
            except Exception as e:  # noqa: BLE001
                if not stop.is_set():
                    errors.append(repr(e))
        return run

    def channel():
        """None while the plugin socket is away (kubelet-restart window; grpcio's shared reconnect backoff can outlast it)."""
        try:
            return d.channel()
        except grpc.FutureTimeoutError:
            return None

    def allocator():
        ch = None
        try:
            while not stop.is_set():
                if ch is None:
                    ch = channel()
                    continue
                try:
                    r = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))[0]
                    results.append(r["ALIYUN_COM_GPU_MEM_IDX"])
                except grpc.RpcError:
                    time.sleep(0.05)
                    ch.close()
                    ch = None
        finally:
            if ch is not None:
                ch.close()

    def churn():
        with lock:
            i = next_id[0]
            next_id[0] += 1
        world.kube.add_pod(make_pod(i, NODE, gpu_mem=2, idx=i % 8, assume_time=1_800_000_000_000_000_000 + i))
        if i % 3 == 0:
            try:
                world.kube.delete_pod(f"pod-{i - 2}")
            except KeyError:
                pass
        time.sleep(0.004)

    def watcher():
        ch = channel()
        if ch is None:
            return
        try:
            call = d.kubelet.list_and_watch(ch)
            next(iter(call))
            time.sleep(0.05)
            call.cancel()
        except grpc.RpcError:
            time.sleep(0.05)
        finally:
            ch.close()

    def health():
        ch = channel()
        if ch is None:
            return
        try:
            d.inject(ch, fakes.UUIDS[3], 8, 31)   # benign: changes nothing
            d.inject(ch, fakes.UUIDS[6], 0x100, 1)
        except grpc.RpcError:
            pass
        finally:
            ch.close()
        time.sleep(0.02)

    ts = [threading.Thread(target=guard(f)) for f in (churn, watcher, watcher, health)] + \
         [threading.Thread(target=allocator) for _ in range(4)]
    [t.start() for t in ts]
    half = float(os.environ.get("GSB_SOAK_SECONDS", "4")) / 2  # longer runs: GSB_SOAK_SECONDS=60 tools/sanitize.sh
    time.sleep(half)
    d.kubelet.stop()  # kubelet restart: the daemon rebuilds and registers again while everything keeps going
    time.sleep(0.2)
    d.kubelet.start()
    assert d.kubelet.register_requests.get(timeout=20) == d.register_request
    time.sleep(half)
    stop.set()
    [t.join(30) for t in ts]
    assert not errors, errors[:3]
    assert d.proc.poll() is None and len(results) > 50
    patched = list(world.kube.patched_ok)  # applied PATCHes only: a pod deleted by the churn answers 404, possibly twice
    twice = sorted({p for p in patched if patched.count(p) > 1})
    if twice:  # show what the daemon logged about the first such pod
        name = twice[0].rsplit("/", 1)[1]
        d.log.flush()
        print("\n".join(l for l in open(world.dir / "gsbd.log").read().splitlines() if name in l or "Failed" in l)[-3000:])
    assert not twice  # no pod handed out twice
    # the calls in flight when the kubelet restarted lost their responses with the old socket
    ok = len([r for r in results if r != "-1"])
    assert len(patched) - 8 <= ok <= len(patched)
    ch = d.channel()
    d.inject(ch, fakes.UUIDS[6], 0x100, 1)  # the rebuilt plugin starts all-Healthy, like the reference's: mark again
    deadline, devs = time.time() + 5, None
    while time.time() < deadline:
        devs = wo.unmarshal_ListAndWatchResponse(next(iter(d.kubelet.list_and_watch(ch))))
        if devs == all_devs(unhealthy={6}):
            break
        time.sleep(0.1)
    assert devs == all_devs(unhealthy={6})
    ch.close()
    d.proc.send_signal(signal.SIGTERM)
    assert d.proc.wait(timeout=20) == 0


def test_resource_quantity_spellings_are_read_like_quantity_value(world):
    """podutils.go:122-131 sums `limits["aliyun.com/gpu-mem"].Value()`: resource.Quantity accepts decimal and binary SI
    suffixes, exponents and fractions, and Value() rounds fractions up. The daemon's reader (C++), the Python front
    end's and the oracle's agree on every spelling below."""
    from gpushare_device_plugin_b200.nvidia import podutils
    spellings = {"2": 2, "2.0": 2, "2e0": 2, "20e-1": 2, "0.002k": 2, "1999m": 2, "2000m": 2, "1500m": 2, "2001m": 3,
                 "2.1": 3, "3": 3, "0.003k": 3, "1Ki": 1024, "1k": 1000, "0": 0, "1e3": 1000, "5E-1": 1}
    for text, want in spellings.items():
        assert wo.quantity_value(text) == want and podutils.quantityValue(text) == want, text
    base = 1_500_000_000_000_000_000  # older than every config-4 pod: these are picked first, by age
    expect = []
    for i, (text, want) in enumerate(spellings.items()):
        p = make_pod(300 + i, NODE, gpu_mem=1, idx=i % 8, assume_time=base + i)
        p["spec"]["containers"][0]["resources"]["limits"]["aliyun.com/gpu-mem"] = text
        world.kube.add_pod(p)
        expect.append((want, str(i % 8)))
    d = world.start("--pod-informer=false", "--pod-cache-ttl", "0")
    ch = d.channel()
    for size in (2, 3):
        req = wo.marshal_AllocateRequest([["x"] * size])
        for want, idx in [e for e in expect if e[0] == size]:
            assert wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == idx
        # none left of that size among the odd spellings
        got = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"]
        assert got == "-1"
    ch.close()


def test_register_waits_for_a_kubelet_that_is_still_coming_up(world, tmp_path):
    """server.go:90-104 dials with grpc.WithBlock() and a 5 s timeout: a kubelet whose socket appears a moment later
    is waited for, not treated as a start-up failure (exit 2 only after the deadline, see the exit-code test)."""
    kubelet = FakeKubelet(str(tmp_path))
    kubelet.stop()  # not there yet
    d = world.start(wait_register=False, kubelet=kubelet)
    time.sleep(1.5)
    assert d.proc.poll() is None
    kubelet.start()
    assert kubelet.register_requests.get(timeout=10) == wo.marshal_RegisterRequest("v1beta1", "aliyungpushare.sock", "aliyun.com/gpu-mem")
    ch = d.channel()
    assert len(wo.unmarshal_ListAndWatchResponse(next(iter(kubelet.list_and_watch(ch))))) == 1432
    ch.close()


def test_serialize_allocate_is_the_reference_locking(world):
    """allocate.go:59-60 holds the plugin lock across the whole call, the apiserver PATCH included: N concurrent
    Allocates take N round trips. `--serialize-allocate` reproduces that (the compiled reference-behaviour row of the
    benchmark); the default releases the lock before the PATCH."""
    world.kube.patch_delay = 0.15
    timings = {}
    for mode, flags in (("default", ()), ("serialized", ("--serialize-allocate",))):
        d = world.start("--pod-informer=false", "--pod-cache-ttl", "60", *flags)
        chans = [d.channel() for _ in range(4)]
        wo.unmarshal_AllocateResponse(d.kubelet.allocate(chans[0], wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))  # warm
        out = []
        t0 = time.time()
        ts = [threading.Thread(target=lambda c=c: out.append(wo.unmarshal_AllocateResponse(
            d.kubelet.allocate(c, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))[0]["ALIYUN_COM_GPU_MEM_IDX"])) for c in chans]
        [t.start() for t in ts]
        [t.join(30) for t in ts]
        timings[mode] = time.time() - t0
        assert len(out) == 4 and "-1" not in out
        [c.close() for c in chans]
        d.close()
    assert timings["default"] < 0.45 and timings["serialized"] > 0.55, timings  # ~1 vs ~4 PATCH round trips of 0.15 s


# ---- round-2 advisor findings ---------------------------------------------------------------------------

def test_health_event_raised_before_any_stream_is_in_the_first_frame(world):
    """A fault raised while no ListAndWatch stream is attached (start-up walk, NOT_SUPPORTED registration, kubelet
    reconnect) must be in the next stream's first frame: the producer writes the node state, streams only resend."""
    d = world.start()
    ch = d.channel()
    d.inject(ch, fakes.UUIDS[3], 8, 79)
    time.sleep(0.6)  # the health loop has taken the event; no stream has ever been open
    it = Frames(d.kubelet.list_and_watch(ch))
    assert wo.unmarshal_ListAndWatchResponse(next_frame(it)) == all_devs(unhealthy={3})
    assert next_frame(it, 0.6) == "timeout"
    it2 = Frames(d.kubelet.list_and_watch(ch))  # a reconnect starts from the same state
    assert wo.unmarshal_ListAndWatchResponse(next_frame(it2)) == all_devs(unhealthy={3})
    d.inject(ch, fakes.UUIDS[6], 0x100, 1)
    for s in (it, it2):
        assert wo.unmarshal_ListAndWatchResponse(next_frame(s)) == all_devs(unhealthy={3, 6})
    assert "unhealthy_slices 179" in _dump(d)
    ch.close()


def _dump(d):
    d.proc.send_signal(signal.SIGQUIT)
    deadline = time.time() + 5
    while time.time() < deadline:
        files = [f for f in os.listdir(d.dir) if f.startswith("go_")]
        if files:
            time.sleep(0.1)
            return open(os.path.join(d.dir, files[0])).read()
        time.sleep(0.05)
    return ""


def test_reference_stream_mode_is_exact_for_a_stream_that_attaches_late(world):
    """coalesce-health=false: one resend per fake device, each frame the state after that event — also for the
    events that arrive after a stream attached to a node that already had an unhealthy GPU."""
    d = world.start("--coalesce-health=false")
    ch = d.channel()
    d.inject(ch, fakes.UUIDS[0], 8, 79)
    time.sleep(0.6)
    it = Frames(d.kubelet.list_and_watch(ch))
    start = all_devs(unhealthy={0})
    assert wo.unmarshal_ListAndWatchResponse(next_frame(it)) == start
    d.inject(ch, fakes.UUIDS[1], 8, 79)
    ids = [x[0] for x in start]
    want = wo.list_and_watch_stream(start, wo.xid_event_effects(ids, 8, 79, fakes.UUIDS[1]))
    assert [next_frame(it) for _ in range(179)] == want[1:]
    ch.close()


def test_single_gpu_node_does_not_answer_a_stale_cache_with_the_shortcut(world):
    """One GPU + a cached pod table older than the pod being started: the reference (LIST per call) finds the pod,
    answers with its annotated index and PATCHes it; the single-GPU shortcut must not answer from the stale table."""
    for k in list(world.kube.pods):
        world.kube.pods[k]["status"]["phase"] = "Running"
    d = world.start("--pod-cache-ttl", "60", "--pod-informer=false", n_gpus=1)
    ch = d.channel()
    envs = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c"]])))
    assert envs[0]["NVIDIA_VISIBLE_DEVICES"] == fakes.UUIDS[0]  # nothing pending: the shortcut is right here
    new = make_pod(99, NODE, gpu_mem=2, idx=fakes.MINORS[0], assume_time=1_800_000_000_000_000_000)
    with world.kube.lock:
        world.kube.pods[("default", "pod-99")] = new
        world.kube.order.append(("default", "pod-99"))
    envs = wo.unmarshal_AllocateResponse(d.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))
    assert envs[0]["NVIDIA_VISIBLE_DEVICES"] == str(fakes.MINORS[0]) == envs[0]["ALIYUN_COM_GPU_MEM_IDX"]
    assert world.kube.pod("pod-99")["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true"
    ch.close()


def test_inventory_event_marks_only_that_gpu_unhealthy(world):
    """GSB_EVENT_INVENTORY (0x200): the library's off-path NVML refresh found a GPU whose identity or total no longer
    matches what is advertised. The daemon must not keep advertising it as if nothing happened: every fake device of
    that GPU goes Unhealthy (the others stay), like any other device-scoped fault (nvidia.go:146-150)."""
    d = world.start()
    ch = d.channel()
    it = Frames(d.kubelet.list_and_watch(ch))
    assert next_frame(it) == wo.marshal_ListAndWatchResponse(all_devs())
    d.inject(ch, fakes.UUIDS[4], 0x200, 2)  # total changed
    assert wo.unmarshal_ListAndWatchResponse(next_frame(it)) == all_devs(unhealthy={4})
    d.inject(ch, fakes.UUIDS[4], 0x300, 1)  # an event type nobody knows is ignored (nvidia.go:127-129)
    assert next_frame(it, 0.7) == "timeout"
    ch.close()
