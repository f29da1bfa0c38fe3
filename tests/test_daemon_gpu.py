"""The real daemon process on a real GPU: `python -m gpushare_device_plugin_b200.cmd.nvidia` with the
DaemonSet's command line, against a fake kubelet and a mock apiserver — registration, the 179-slice
list, Allocate, the active HBM probe flagging corruption, re-registration when kubelet.sock is
re-created (gpumanager.go:83-87), clean exit on SIGTERM."""
import os
import signal
import subprocess
import sys
import time

import pytest

from gpushare_device_plugin_b200.testing.fake_kubelet import FakeKubelet
from gpushare_device_plugin_b200.testing.mock_kube import MockKube, make_node, make_pod
from oracle import wire_oracle as wo

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE = "b200-0"


GSBD = os.path.join(ROOT, "gpushare_device_plugin_b200", "gsbd")


# probe memory: a standing 2 GiB arena after a start-up walk of everything allocatable, or the shipped default — no
# standing arena, one transient window per cycle
ARENA = ["--probe-arena-mib", "2048", "--startup-full-walk"]
DEFAULT = []


@pytest.mark.parametrize("front_end,probe_flags", [("python", ARENA), ("native", ARENA), ("native", DEFAULT)],
                         ids=["python-arena", "native-arena", "native-transient-default"])
def test_daemon_end_to_end(tmp_path, front_end, probe_flags):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    uuid0, minor0 = pynvml.nvmlDeviceGetUUID(h), pynvml.nvmlDeviceGetMinorNumber(h)
    n_gpus = pynvml.nvmlDeviceGetCount()
    uuid0 = uuid0.decode() if isinstance(uuid0, bytes) else uuid0
    pods = [make_pod(i, NODE, gpu_mem=4, idx=minor0, assume_time=1_700_000_000_000_000_000 + i) for i in range(3)]
    kube = MockKube(make_node(NODE), pods)
    kubelet = FakeKubelet(str(tmp_path))
    kubeconfig = tmp_path / "kubeconfig"
    kubeconfig.write_text(f"apiVersion: v1\nkind: Config\ncurrent-context: c\nclusters:\n- name: k\n  cluster:\n    server: {kube.url}\n"
                          "contexts:\n- name: c\n  context:\n    cluster: k\n    user: u\nusers:\n- name: u\n  user:\n    token: t\n")
    env = dict(os.environ, KUBECONFIG=str(kubeconfig), NODE_NAME=NODE, GPUSHARE_PLUGIN_DIR=str(tmp_path) + "/",
               GPUSHARE_DUMP_DIR=str(tmp_path), PYTHONPATH=ROOT)
    log = open(tmp_path / "daemon.log", "w")
    argv = [sys.executable, "-m", "gpushare_device_plugin_b200.cmd.nvidia"] if front_end == "python" else [GSBD]
    proc = subprocess.Popen(argv + ["-logtostderr", "--v=5", "--memory-unit=GiB", "--health-check", "--token", "t",
                                    "--probe-period-ms", "100"] + probe_flags,
                            env=env, stderr=log, stdout=log, cwd=ROOT)
    try:
        req = kubelet.register_requests.get(timeout=120)
        assert req == wo.marshal_RegisterRequest("v1beta1", "aliyungpushare.sock", "aliyun.com/gpu-mem")
        assert kube.nodes[NODE]["status"]["capacity"]["aliyun.com/gpu-count"] == str(n_gpus)
        ch = kubelet.channel("aliyungpushare.sock")
        stream = kubelet.list_and_watch(ch)
        devs = wo.unmarshal_ListAndWatchResponse(next(stream))
        assert len(devs) == 179 * n_gpus and devs[0] == [uuid0 + "-_-0", "Healthy"]
        envs = wo.unmarshal_AllocateResponse(kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))
        assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == str(minor0) and envs[0]["ALIYUN_COM_GPU_MEM_DEV"] == "179"
        assert kube.pod("pod-00")["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true"
        # the prober has been rotating clean windows for a while: nothing was re-sent
        time.sleep(1.5)
        if not probe_flags:  # transient default: between cycles the plugin holds no HBM beyond its CUDA context
            free = [pynvml.nvmlDeviceGetMemoryInfo(h).free for _ in range(5) if time.sleep(0.03) is None]
            assert max(free) > pynvml.nvmlDeviceGetMemoryInfo(h).total - 3 * (1 << 30), free
        # SIGQUIT -> thread dump file (gpumanager.go:97-101), daemon keeps running
        proc.send_signal(signal.SIGQUIT)
        deadline = time.time() + 10
        while time.time() < deadline and not [f for f in os.listdir(tmp_path) if f.startswith("go_")]:
            time.sleep(0.1)
        assert [f for f in os.listdir(tmp_path) if f.startswith("go_")] and proc.poll() is None
        # kubelet restart: kubelet.sock re-created -> plugin rebuilds and registers again
        ch.close()
        kubelet.stop()
        time.sleep(0.3)
        kubelet.start()
        req2 = kubelet.register_requests.get(timeout=60)
        assert req2 == req
        ch = kubelet.channel("aliyungpushare.sock")
        assert len(wo.unmarshal_ListAndWatchResponse(next(kubelet.list_and_watch(ch)))) == 179 * n_gpus
        ch.close()
        proc.send_signal(signal.SIGTERM)
        assert proc.wait(timeout=30) == 0
        assert not os.path.exists(tmp_path / "aliyungpushare.sock")
    finally:
        if proc.poll() is None:
            proc.kill()
        log.close()
        kubelet.stop()
        kube.close()
        print(open(tmp_path / "daemon.log").read()[-3000:])
