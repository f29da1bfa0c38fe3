"""tests/golden/allocate_cases.json — the frozen answers of the host-logic oracle (oracle/make_wire_golden.py) — held
against BOTH the oracle (drift in the restatement shows up here) and the product (gsb_allocate through the C ABI).
PARITY UNPINNED: the file holds the oracle's reading of allocate.go / podutils.go / podmanager.go, not outputs of
the reference, which has no tests on this path and cannot run here; each case cites the lines it exercises."""
import copy
import ctypes as C
import json
import os
import random
import subprocess
import sys

import pytest

from gpushare_device_plugin_b200 import _abi
from gpushare_device_plugin_b200.nvidia.allocate import AllocateContext, pod_table
from oracle import wire_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "allocate_cases.json")))
KIND = {_abi.GSB_ALLOC_MATCHED: "matched", _abi.GSB_ALLOC_SINGLE_GPU: "single_gpu", _abi.GSB_ALLOC_ERR_RESPONSE: "err_response"}


def test_the_committed_file_is_what_the_generator_writes():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_wire_golden.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: c["name"][:60])
def test_oracle_still_gives_the_frozen_answer(case):
    ctx = case["ctx"]
    envs, matched = wo.Allocate(case["container_requests"], copy.deepcopy(case["pods"]), ctx["node"], ctx["devNameMap"],
                                ctx["slices"], ctx["unit"], ctx["disable_cgpu_isolation"])
    want = case["want"]
    assert envs == want["envs"] and (matched["metadata"]["uid"] if matched else None) == want["matched_uid"]
    assert wo.marshal_AllocateResponse(envs).hex() == want["response_hex"]


@pytest.mark.parametrize("pods_unique", [False, True], ids=["dedupe", "uid-keyed-table"])
@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: c["name"][:60])
def test_product_gives_the_frozen_answer(case, pods_unique):
    ctx, want = case["ctx"], case["want"]
    uids = [p["metadata"].get("uid") for p in case["pods"]]
    if pods_unique and len(set(uids)) != len(uids):
        pytest.skip("a uid-keyed table cannot hold duplicate UIDs")
    actx = AllocateContext(ctx["devNameMap"], ctx["slices"], ctx["unit"] == "GiB", ctx["disable_cgpu_isolation"])
    actx.ctx.pods_unique = 1 if pods_unique else 0
    req = wo.marshal_AllocateRequest(case["container_requests"])
    table, _keep = pod_table(case["pods"], ctx["node"])
    buf = C.create_string_buffer(1 << 18)
    n, pidx, preq = C.c_size_t(0), C.c_int32(-1), C.c_uint32(0)
    kind = _abi.lib.gsb_allocate(C.byref(actx.ctx), table, len(case["pods"]), req, len(req), buf, len(buf), C.byref(n),
                                 C.byref(pidx), C.byref(preq))
    assert kind > 0, _abi.last_error()
    assert KIND[kind] == want["kind"] and preq.value == want["pod_req_gpu"]
    assert buf.raw[: n.value].hex() == want["response_hex"]  # sorted-key entry order: one of gogo's possible orders
    assert wo.unmarshal_AllocateResponse(buf.raw[: n.value]) == want["envs"]
    got_uid = case["pods"][pidx.value]["metadata"]["uid"] if pidx.value >= 0 else None
    assert got_uid == want["matched_uid"]


def test_tie_order_of_go110_sort_product_vs_oracle_randomised():
    """The order among EQUAL assume-times is decided by Go 1.10's sort.Sort run with the reference's non-strict Less
    (podmanager.go:256-258); product (C++) and oracle (Python) restate it independently: up to 300 candidates, few
    distinct times, several sizes — the first matching pod must be the same one."""
    rng = random.Random(20260921)
    node = "b200-0"
    minors = {f"GPU-{i:08x}-4820-abfc-e83e-943181975700": i for i in range(8)}
    for trial in range(300):
        n = rng.choice([13, 14, 20, 40, 41, 42, 64, 100, 300]) if trial % 3 else rng.randrange(1, 60)
        times = rng.choice([1, 2, 3, 8])
        sizes = rng.choice([1, 2, 3])
        pods = []
        for i in range(n):
            pods.append({"metadata": {"name": f"p{i}", "namespace": "d", "uid": f"u{i}",
                                      "annotations": {wo.EnvResourceIndex: str(i % 8), wo.EnvAssignedFlag: "false",
                                                      wo.EnvResourceAssumeTime: str(1000 + rng.randrange(times))}},
                         "spec": {"nodeName": node, "containers": [{"resources": {"limits": {wo.resourceName: str(1 + rng.randrange(sizes))}}}]}})
        want_req = 1 + rng.randrange(sizes)
        reqs = [[f"x{j}" for j in range(want_req)]]
        envs, matched = wo.Allocate(reqs, copy.deepcopy(pods), node, minors, 179, "GiB", False)
        actx = AllocateContext(minors, 179, True, False)
        req = wo.marshal_AllocateRequest(reqs)
        table, _keep = pod_table(pods, node)
        buf = C.create_string_buffer(1 << 16)
        nn, pidx, preq = C.c_size_t(0), C.c_int32(-1), C.c_uint32(0)
        kind = _abi.lib.gsb_allocate(C.byref(actx.ctx), table, len(pods), req, len(req), buf, len(buf), C.byref(nn),
                                     C.byref(pidx), C.byref(preq))
        assert kind > 0
        got = pods[pidx.value]["metadata"]["uid"] if pidx.value >= 0 else None
        assert got == (matched["metadata"]["uid"] if matched else None), (trial, n, times, sizes)
        assert wo.unmarshal_AllocateResponse(buf.raw[: nn.value]) == envs


# ---- the same frozen answers through the DAEMONS, end to end -------------------------------------------------------
# gsbd turns the apiserver's pod JSON into the gsb_pod table with its own C++ code (csrc/daemon/gsbd_pods.hpp: Atoi,
# ParseUint, quantity values, node filter) — a third implementation of podutils.go next to the oracle's and the Python
# front end's. Every frozen case that fits the synthetic 8-GPU node (default ctx: 179 GiB slices, minors 2,3,0,1,6,7,4,5)
# is replayed through a real gsbd: mock apiserver holding the case's pods -> gRPC Allocate -> decoded envs and the
# annotation PATCH on exactly the pod the oracle matched.

def _daemon_cases():
    from tests import fakes
    want_map = dict(zip(fakes.UUIDS, fakes.MINORS))
    return [c for c in GOLD["cases"]
            if c["ctx"]["devNameMap"] == want_map and c["ctx"]["slices"] == 179 and c["ctx"]["unit"] == "GiB"
            and not c["ctx"]["disable_cgpu_isolation"]]


def _load_case(kube, case, tag=""):
    """`tag` makes the UIDs of this case unique for the daemon's lifetime: a pod this daemon handed out is never a
    candidate again under the same UID (its double-hand-out guard), and the fixture reuses uid-0, uid-1 ... per case."""
    with kube.lock:
        kube.pods.clear()
        kube.order.clear()
        for i, p in enumerate(copy.deepcopy(case["pods"])):
            if "uid" in p["metadata"]:
                p["metadata"]["uid"] = tag + p["metadata"]["uid"]
            key = (p["metadata"].get("namespace", "default"), p["metadata"]["name"])
            if key in kube.pods:  # duplicate-UID cases reuse nothing else; names are unique in the fixture
                key = (key[0], key[1] + "-%d" % i)
            kube.pods[key] = p
            kube.order.append(key)


def _check_case(kube, case, envs):
    want = case["want"]
    assert envs == want["envs"], case["name"]
    patched = sorted(k[1] for k, p in kube.pods.items()
                     if (p["metadata"].get("annotations") or {}).get(wo.EnvAssignedFlag) == "true"
                     and not any((q["metadata"].get("annotations") or {}).get(wo.EnvAssignedFlag) == "true"
                                 and q["metadata"]["name"] == p["metadata"]["name"] for q in case["pods"]))
    if want["matched_uid"] is None:
        assert patched == [], (case["name"], patched)
    else:
        matched_name = next(p["metadata"]["name"] for p in case["pods"] if p["metadata"]["uid"] == want["matched_uid"])
        assert patched == [matched_name], (case["name"], patched, matched_name)


def test_native_daemon_gives_the_frozen_answers_end_to_end(tmp_path):
    import signal
    from gpushare_device_plugin_b200.testing.fake_kubelet import FakeKubelet
    from gpushare_device_plugin_b200.testing.mock_kube import MockKube, make_node
    gsbd = os.environ.get("GSBD_BINARY") or os.path.join(ROOT, "gpushare_device_plugin_b200", "gsbd")
    if not os.access(gsbd, os.X_OK):
        pytest.skip("gsbd not built")
    cases = _daemon_cases()
    assert len(cases) >= 45
    kube = MockKube(make_node("b200-0"), [])
    kubelet = FakeKubelet(str(tmp_path))
    env = dict(os.environ, NODE_NAME="b200-0", GPUSHARE_PLUGIN_DIR=str(tmp_path) + "/", GPUSHARE_RETRY_SLEEP_MS="1",
               GSBD_ALLOW_FAKE_INVENTORY="1")
    env.pop("KUBECONFIG", None)
    log = open(tmp_path / "gsbd.log", "w")
    # LIST per call (the reference's behaviour): the table is rebuilt from the case's pods on every request
    proc = subprocess.Popen([gsbd, "--v=5", "--fake-inventory", "8", "--kube-api-url", kube.url, "--pod-informer=false",
                             "--pod-cache-ttl", "0"], env=env, stderr=log, stdout=log)
    try:
        kubelet.register_requests.get(timeout=20)
        ch = kubelet.channel("aliyungpushare.sock")
        for n, case in enumerate(cases):
            _load_case(kube, case, "c%d-" % n)
            raw = kubelet.allocate(ch, wo.marshal_AllocateRequest(case["container_requests"]))
            _check_case(kube, case, wo.unmarshal_AllocateResponse(raw))
            assert raw.hex() == case["want"]["response_hex"]
        ch.close()
    finally:
        proc.send_signal(signal.SIGTERM)
        proc.wait(timeout=10)
        log.close()
        kubelet.stop()
        kube.close()


def test_python_front_end_gives_the_frozen_answers_end_to_end(tmp_path, monkeypatch):
    import time
    from gpushare_device_plugin_b200.nvidia import kubeclient, podmanager, server
    from gpushare_device_plugin_b200.testing.fake_kubelet import FakeKubelet
    from gpushare_device_plugin_b200.testing.mock_kube import MockKube, make_node
    from tests import fakes
    fakes.install(monkeypatch)
    monkeypatch.setattr(time, "sleep", lambda s: None)
    kube = MockKube(make_node("b200-0", labels={}), [])
    podmanager.kubeInit(kubeclient.Clientset(kube.url), "b200-0")
    kubelet = FakeKubelet(str(tmp_path))
    p = server.NewNvidiaDevicePlugin(False, False, False, None, socket=str(tmp_path / "aliyungpushare.sock"), pod_cache_ttl=0)
    try:
        p.Serve(kubelet.socket)
        ch = kubelet.channel("aliyungpushare.sock")
        for n, case in enumerate(_daemon_cases()):
            _load_case(kube, case, "c%d-" % n)
            raw = kubelet.allocate(ch, wo.marshal_AllocateRequest(case["container_requests"]))
            _check_case(kube, case, wo.unmarshal_AllocateResponse(raw))
        ch.close()
    finally:
        p.Stop()
        kubelet.stop()
        kube.close()
