"""cmd/inspect parity smoke (BASELINE.json config[0]): synthetic cluster behind
the mock APISERVER (the reference's inspect talks to the apiserver through $KUBECONFIG, not to the
kubelet), expected text derived by hand from display.go + text/tabwriter's rules."""
import pytest

from gpushare_device_plugin_b200.cmd import inspect as ins
from gpushare_device_plugin_b200.cmd.tabwriter import Writer
from gpushare_device_plugin_b200.nvidia import kubeclient
from gpushare_device_plugin_b200.testing.mock_kube import MockKube, make_node, make_pod

NODE = "b200-0"


def test_tabwriter_elastic_tabstops():
    w = Writer(0, 0, 2, " ", 0)
    w.write("a\tb\tc\n")
    w.write("aa\tbb\tc\n")
    w.write("aaa\t\n")   # column 0 block continues, column 1 block ended (its cell is the trailing one)
    w.write("x\ty\n")
    assert w.flush() == "a    b   c\naa   bb  c\naaa  \nx    y\n"
    w.write("no tabs here\n")
    w.write("k\tv\n")
    assert w.flush() == "no tabs here\nk  v\n"
    w.write("1\t22\t\n333\t4\t\n")  # trailing tab: last column is padded too
    assert w.flush() == "1    22  \n333  4   \n"


@pytest.fixture
def cluster():
    node = make_node(NODE, gpu_count=8)
    node["status"]["allocatable"]["aliyun.com/gpu-mem"] = str(8 * 179)
    node["status"]["capacity"]["aliyun.com/gpu-mem"] = str(8 * 179)
    node["status"]["addresses"] = [{"type": "Hostname", "address": NODE}, {"type": "InternalIP", "address": "10.0.0.7"}]
    pods = [make_pod(i, NODE, gpu_mem=4, idx=i % 8, assume_time=1, assigned="true", phase="Running") for i in range(64)]
    pods.append(make_pod(64, NODE, gpu_mem=4, idx=0, assume_time=1, assigned="true", phase="Succeeded"))  # not active
    pods.append(make_pod(65, "other-node", gpu_mem=4, idx=0, assume_time=1, phase="Running"))
    kube = MockKube(node, pods)
    other = make_node("cpu-only")
    kube.nodes["cpu-only"] = other  # no gpu-mem: filtered by isGPUSharingNode
    yield kube
    kube.close()


def test_summary_config1(cluster):
    out = ins.run([], kubeclient.Clientset(cluster.url))
    lines = out.split("\n")
    # unit heuristic reproduced: 179 > 100 => "MiB" (nodeinfo.go:238-242)
    assert lines[0].split("  ")[-1].strip() == "GPU Memory(MiB)" or lines[0].endswith("GPU Memory(MiB)")
    hdr_cols = lines[0].split()
    assert hdr_cols[:3] == ["NAME", "IPADDRESS", "GPU0(Allocated/Total)"] and hdr_cols[9] == "GPU7(Allocated/Total)"
    row = lines[1].split()
    assert row == [NODE, "10.0.0.7"] + ["32/179"] * 8 + ["256/1432"]
    # header and row cells start at the same offsets (elastic tabstops, padding 2)
    for k in range(8):
        assert lines[0].index(f"GPU{k}(Allocated/Total)") == [i for i in range(len(lines[1])) if lines[1].startswith("32/179", i)][k]
    first_row_raw = f"{NODE}\t10.0.0.7\t" + "32/179\t" * 8 + "256/1432\n"
    assert lines[2] == "-" * (len(first_row_raw) + 20)   # prtLineLen = buf.Len() + 20 (display.go:216-218)
    assert lines[3] == "Allocated/Total GPU Memory In Cluster:"
    assert lines[4] == "256/1432 (17%)  " and out.endswith("\n")


def test_single_node_details_and_pending_column(cluster):
    # a pod without IDX annotation lands in the pending pseudo-device -1
    with cluster.lock:
        p = make_pod(70, NODE, gpu_mem=6, idx=None, assume_time=None, assigned=None, phase="Pending")
        cluster.pods[("default", "pod-70")] = p
        cluster.order.append(("default", "pod-70"))
        # scheduler-framework allocation annotation wins over IDX (nodeinfo.go:244-271)
        q = make_pod(71, NODE, gpu_mem=5, idx=3, assume_time=1, phase="Running")
        q["metadata"]["annotations"]["scheduler.framework.gpushare.allocation"] = '{"0":{"1":2,"2":3}}'
        cluster.pods[("default", "pod-71")] = q
        cluster.order.append(("default", "pod-71"))
    out = ins.run([NODE], kubeclient.Clientset(cluster.url))
    row = out.split("\n")[1].split()
    assert row == [NODE, "10.0.0.7", "32/179", "34/179", "35/179"] + ["32/179"] * 5 + ["6", "267/1432"]
    assert "PENDING(Allocated)" in out.split("\n")[0]
    det = ins.run(["-d", NODE], kubeclient.Clientset(cluster.url))
    assert f"NAME:       {NODE}" in det and "IPADDRESS:  10.0.0.7" in det
    rows = {ln.split()[0]: ln.split() for ln in det.split("\n") if ln.startswith("pod-")}
    assert len(rows) == 66  # 64 + pending + framework pod, each listed once (exists[pod.UID])
    assert rows["pod-09"][2:] == ["0", "4", "0", "0", "0", "0", "0", "0", "0"]      # IDX 1, plus the pending column
    assert rows["pod-70"][2:] == ["0"] * 8 + ["6"]
    assert rows["pod-71"][2:] == ["0", "2", "3", "0", "0", "0", "0", "0", "0"]
    assert "Allocated :  267 (18%)" in det and "Total :      1432" in det
    assert det.endswith("Allocated/Total GPU Memory In Cluster:  267/1432 (18%)  \n")  # trailing tab => padded cell


def test_output_text_matches_the_frozen_fixture():
    """tests/golden/inspect_cases.json (tests/golden/make_inspect_golden.py): regression freeze of the whole output
    text — parity unpinned, see the generator's header."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "golden", "make_inspect_golden.py"), "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
