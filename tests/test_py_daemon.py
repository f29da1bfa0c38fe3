"""The Python daemon's lifecycle on CPU (cmd/nvidia.py + nvidia/gpumanager.py + watchers.py + coredump.py): the same
scenario tests/test_daemon_gpu.py runs on a B200, with the synthetic inventory injected by tests/helpers/py_daemon.py.
Reference behaviour: gpumanager.go:33-111 (restart on kubelet.sock CREATE and SIGHUP, dump on SIGQUIT, clean stop on
SIGTERM/SIGINT, exit 2 when the plugin cannot start)."""
import os
import signal
import subprocess
import sys
import time

import pytest

from gpushare_device_plugin_b200.testing.fake_kubelet import FakeKubelet
from gpushare_device_plugin_b200.testing.mock_kube import MockKube, config4_pods, make_node
from oracle import wire_oracle as wo

from . import fakes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE = "b200-0"


def start(tmp_path, kube, *extra, n_gpus=8):
    env = dict(os.environ, NODE_NAME=NODE, GPUSHARE_PLUGIN_DIR=str(tmp_path) + "/", GPUSHARE_DUMP_DIR=str(tmp_path),
               PYTHONPATH=ROOT, PY_DAEMON_FAKE_GPUS=str(n_gpus))
    kc = tmp_path / "kubeconfig"
    kc.write_text(f"apiVersion: v1\nkind: Config\ncurrent-context: c\nclusters:\n- name: k\n  cluster:\n    server: {kube.url}\n"
                  "contexts:\n- name: c\n  context:\n    cluster: k\n    user: u\nusers:\n- name: u\n  user:\n    token: t\n")
    env["KUBECONFIG"] = str(kc)
    log = open(tmp_path / "daemon.log", "w")
    p = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "helpers", "py_daemon.py"), "-logtostderr", "--v=5",
                          "--memory-unit=GiB", "--health-check", "--token", "t", *extra], env=env, stderr=log, stdout=log, cwd=ROOT)
    return p, log


@pytest.fixture
def world(tmp_path):
    kube = MockKube(make_node(NODE), config4_pods(NODE))
    kubelet = FakeKubelet(str(tmp_path))
    procs = []

    def run(*extra, **kw):
        p, log = start(tmp_path, kube, *extra, **kw)
        procs.append((p, log))
        return p
    yield type("W", (), {"kube": kube, "kubelet": kubelet, "run": staticmethod(run), "dir": tmp_path})
    for p, log in procs:
        if p.poll() is None:
            p.kill()
        log.close()
    kubelet.stop()
    kube.close()
    print(open(tmp_path / "daemon.log").read()[-2500:])


def test_lifecycle_restart_dump_and_clean_stop(world):
    p = world.run()
    req = world.kubelet.register_requests.get(timeout=60)
    assert req == wo.marshal_RegisterRequest("v1beta1", "aliyungpushare.sock", "aliyun.com/gpu-mem")
    assert world.kube.nodes[NODE]["status"]["capacity"]["aliyun.com/gpu-count"] == "8"  # patchGPUCount (podmanager.go:74-99)
    ch = world.kubelet.channel("aliyungpushare.sock")
    devs = wo.unmarshal_ListAndWatchResponse(next(iter(world.kubelet.list_and_watch(ch))))
    assert len(devs) == 8 * 179 and devs[0] == [fakes.UUIDS[0] + "-_-0", "Healthy"]
    envs = wo.unmarshal_AllocateResponse(world.kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))
    assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0" and envs[0]["ALIYUN_COM_GPU_MEM_DEV"] == "179"
    assert world.kube.pod("pod-00")["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true"
    ch.close()
    p.send_signal(signal.SIGQUIT)  # stack dump to <dumpDir>/go_<ts>.txt, keep running (gpumanager.go:97-101)
    def dumped():
        return [f for f in os.listdir(world.dir) if f.startswith("go_") and os.path.getsize(world.dir / f) > 0]
    deadline = time.time() + 10
    while time.time() < deadline and not dumped():
        time.sleep(0.1)
    assert dumped() and p.poll() is None
    world.kubelet.stop()  # kubelet restart: kubelet.sock re-created -> rebuild + re-register (gpumanager.go:83-87)
    time.sleep(0.3)
    world.kubelet.start()
    assert world.kubelet.register_requests.get(timeout=60) == req
    p.send_signal(signal.SIGHUP)  # gpumanager.go:94-96
    assert world.kubelet.register_requests.get(timeout=60) == req
    ch = world.kubelet.channel("aliyungpushare.sock")
    assert len(wo.unmarshal_ListAndWatchResponse(next(iter(world.kubelet.list_and_watch(ch))))) == 8 * 179
    ch.close()
    p.send_signal(signal.SIGTERM)
    assert p.wait(timeout=30) == 0 and not os.path.exists(world.dir / "aliyungpushare.sock")
    log = open(world.dir / "daemon.log").read()
    assert "inotify:" in log and "Received SIGHUP, restarting." in log and 'Received signal "SIGTERM", shutting down.' in log


def test_exit_code_2_when_the_kubelet_is_not_there(world):
    world.kubelet.stop()  # no kubelet.sock to register with: Serve fails, exit 2 (gpumanager.go:76)
    p = world.run()
    assert p.wait(timeout=60) == 2
    assert "Failed to start device plugin due to" in open(world.dir / "daemon.log").read()


def test_exit_code_1_when_the_node_cannot_be_read(tmp_path):
    kube = MockKube(make_node("some-other-node"), [])  # NODE_NAME is unknown to the apiserver: patchGPUCount fails
    kubelet = FakeKubelet(str(tmp_path))
    p, log = start(tmp_path, kube)
    try:
        assert p.wait(timeout=60) == 1  # gpumanager.go:73
        assert "Failed to get device plugin due to" in open(tmp_path / "daemon.log").read()
    finally:
        if p.poll() is None:
            p.kill()
        log.close()
        kubelet.stop()
        kube.close()
