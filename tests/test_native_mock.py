"""The compiled apiserver stand-in used by bench.py's Allocate legs (csrc/daemon/mock_kube.cc) against the
Python mock the rest of the suite uses (testing/mock_kube.py): same world, same answers, request by request.
Keeps the benchmark harness honest — a plugin that passes against one passes against the other."""
import http.client
import json
import os
import subprocess
import threading

import pytest

from gpushare_device_plugin_b200.testing.mock_kube import MockKube, config4_pods, make_node

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "gpushare_device_plugin_b200", "gsb_mock_kube")
NODE = "b200-0"
SEL = "fieldSelector=spec.nodeName%3Db200-0%2Cstatus.phase%3DPending"


@pytest.fixture
def pair():
    if not os.access(NATIVE, os.X_OK):
        pytest.skip("gsb_mock_kube not built")
    py = MockKube(make_node(NODE, gpu_count=8), config4_pods(NODE, 16))
    proc = subprocess.Popen([NATIVE, "--node", NODE, "--pods", "16"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
    port = int(proc.stdout.readline())
    yield py.port, port
    proc.stdin.close()
    assert proc.wait(timeout=10) == 0
    py.close()


def call(port, method, path, body=None):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=5)
    c.request(method, path, body=body, headers={"Content-Type": "application/strategic-merge-patch+json"} if body else {})
    r = c.getresponse()
    data = r.read()
    c.close()
    return r.status, json.loads(data)


def test_same_answers_request_by_request(pair):
    script = [
        ("GET", "/api/v1/pods?" + SEL, None),
        ("GET", "/api/v1/pods?fieldSelector=spec.nodeName%3Delsewhere", None),
        ("GET", "/api/v1/nodes/" + NODE, None),
        ("GET", "/api/v1/nodes/absent", None),
        ("GET", "/api/v1/nodes", None),
        ("GET", "/pods/", None),
        ("PATCH", "/api/v1/namespaces/default/pods/pod-03",
         '{"metadata":{"annotations":{"ALIYUN_COM_GPU_MEM_ASSIGNED":"true","ALIYUN_COM_GPU_MEM_ASSUME_TIME":"17"}}}'),
        ("PATCH", "/api/v1/namespaces/default/pods/absent", '{"metadata":{"annotations":{"a":"b"}}}'),
        ("PATCH", "/api/v1/namespaces/default/pods/pod-04", "{not json"),
        ("PATCH", "/api/v1/nodes/" + NODE + "/status", '{"status":{"capacity":{"aliyun.com/gpu-count":"4"},"allocatable":{"aliyun.com/gpu-count":"4"}}}'),
        ("GET", "/api/v1/nodes/" + NODE, None),
        ("GET", "/api/v1/pods?" + SEL, None),
        ("GET", "/nothing/here", None),
    ]
    for method, path, body in script:
        a, b = call(pair[0], method, path, body), call(pair[1], method, path, body)
        assert a == b, (method, path, a, b)


def read_events(port, since, n, out):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=10)
    c.request("GET", f"/api/v1/pods?watch=true&{SEL}&resourceVersion={since}")
    r = c.getresponse()
    assert r.status == 200
    for _ in range(n):
        out.append(json.loads(r.readline()))
    c.close()


def test_same_watch_stream(pair):
    streams = []
    for port in pair:
        _, lst = call(port, "GET", "/api/v1/pods?" + SEL)
        rv = int(lst["metadata"]["resourceVersion"])
        got = []
        t = threading.Thread(target=read_events, args=(port, rv, 2, got))
        t.start()
        call(port, "PATCH", "/api/v1/namespaces/default/pods/pod-01", '{"metadata":{"annotations":{"ALIYUN_COM_GPU_MEM_ASSIGNED":"true"}}}')
        call(port, "PATCH", "/api/v1/namespaces/default/pods/pod-02", '{"metadata":{"annotations":{"x":"y"}}}')
        t.join(10)
        assert not t.is_alive() and len(got) == 2
        # a watch opened from the LIST's resourceVersion replays nothing older, one from 0 replays both events
        replay = []
        read_events(port, 0, 2, replay)
        assert replay == got
        streams.append(got)
    assert streams[0] == streams[1]
    assert [e["type"] for e in streams[0]] == ["MODIFIED", "MODIFIED"]
    assert streams[0][0]["object"]["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true"


@pytest.mark.parametrize("impl", ["ours", "ours_py", "reference"])
def test_bench_allocate_leg_runs_clean(impl):
    """bench.py's Allocate leg (quick sizes) for every arm: every request answered, none with the poison envs,
    and the three arms are driven by the same load generator against the same mock."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    r = bench.bench_allocate(impl, quick=True)
    assert r["config4"]["error_responses"] == 0 and r["config4"]["p50_us"] > 0
    assert [s["concurrency"] for s in r["sweep"]] == [1, 16] and all(s["error_responses"] == 0 for s in r["sweep"])
    assert "compiled apiserver stand-in" in r["mock"] and "native HTTP/2 client" in r["client"]
    if impl == "ours":
        src = r["config4_by_pod_source"]
        assert len(src) == 4 and all(v > 0 for v in src.values())
