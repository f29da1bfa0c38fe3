"""cmd/inspect: the product (gpushare_device_plugin_b200/cmd/inspect.py + cmd/tabwriter.py) against an independent
second reading of the reference (oracle/inspect_oracle.py: nodeinfo.go, display.go, podinfo.go, Go's text/tabwriter
restated non-recursively), on the frozen fixture and on randomised clusters. Both are restatements — parity unpinned
(no Go here) — but they were written apart, so a misreading has to be made twice to survive."""
import json
import os

from hypothesis import given, settings, strategies as st

from gpushare_device_plugin_b200.cmd import inspect as ins
from gpushare_device_plugin_b200.cmd.tabwriter import Writer
from oracle import inspect_oracle as io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def product(all_nodes, all_pods, node_name="", details=False):
    """cmd/inspect.run() over plain lists (what it does with the apiserver's answers)."""
    ins.memoryUnit = ""
    if node_name == "":
        nodes = [n for n in all_nodes
                 if ins.quantityValue(((n.get("status") or {}).get("allocatable") or {}).get(ins.resourceName, 0)) > 0]
        pods = ins.filterActivePods(all_pods)
    else:
        nodes = [n for n in all_nodes if n["metadata"]["name"] == node_name]
        pods = ins.filterActivePods([p for p in all_pods if (p.get("spec") or {}).get("nodeName") == node_name])
    infos = ins.buildAllNodeInfos(pods, nodes)
    return ins.displayDetails(infos) if details else ins.displaySummary(infos)


# ---- tabwriter: product's recursive restatement vs the oracle's run-based one ------------------------------------
cell = st.text(alphabet="ab0/() :-%", max_size=7)
line = st.lists(cell, min_size=1, max_size=6).map(lambda cs: "\t".join(cs))


@settings(max_examples=300, deadline=None)
@given(st.lists(line, min_size=0, max_size=12))
def test_tabwriter_two_restatements_agree(lines):
    text = "".join(ln + "\n" for ln in lines)
    w = Writer(0, 0, 2, " ", 0)
    w.write(text)
    assert w.flush() == io.tabwrite(text)


# ---- whole program on random clusters ------------------------------------------------------------------------------
def node(name, n_gpus, per_gpu, ip=True):
    alloc = {}
    if n_gpus is not None:
        alloc["aliyun.com/gpu-count"] = str(n_gpus)
    if per_gpu is not None:
        alloc["aliyun.com/gpu-mem"] = str((n_gpus or 0) * per_gpu if n_gpus else per_gpu)
    addrs = [{"type": "Hostname", "address": name}] + ([{"type": "InternalIP", "address": "10.0.0.%d" % (len(name) % 250)}] if ip else [])
    return {"metadata": {"name": name}, "status": {"allocatable": alloc, "addresses": addrs}}


ALLOCS = [None, '{"0":{"1":2,"2":3}}', '{"0":{"0":4},"1":{"0":1,"3":2}}', '{"0":{"x":2}}', '{"a":{"1":2}}', '{"0":{"1":2.5}}',
          '{"0":{"1":2.0}}', 'not json', '{}', '{"0":null}', '{"0":{"1":null}}', '[1]', 'null', '{"0":{"+1":7}}', '{"0":{"-1":7}}']
IDXS = [None, "0", "1", "3", "7", "-1", "x", "", "+2", "12"]
LIMITS = ["1", "4", "0", "1500m", "2k", None]


@st.composite
def clusters(draw):
    names = draw(st.lists(st.sampled_from(["n-a", "n-b", "gpu-node-long-name", "z"]), min_size=1, max_size=3, unique=True))
    nodes = []
    for nm in names:
        nodes.append(node(nm, draw(st.sampled_from([None, 0, 1, 2, 8])), draw(st.sampled_from([None, 0, 16, 100, 101, 179])),
                          ip=draw(st.booleans())))
    pods = []
    for i in range(draw(st.integers(0, 14))):
        ann = {}
        idx = draw(st.sampled_from(IDXS))
        if idx is not None:
            ann["ALIYUN_COM_GPU_MEM_IDX"] = idx
        alloc = draw(st.sampled_from(ALLOCS))
        if alloc is not None:
            ann["scheduler.framework.gpushare.allocation"] = alloc
        md = {"name": "pod-%d" % i, "namespace": draw(st.sampled_from(["default", "kube-system"])),
              "uid": draw(st.sampled_from(["u%d" % i, "dup"]))}
        if ann or draw(st.booleans()):
            md["annotations"] = ann
        containers = []
        for _ in range(draw(st.integers(0, 2))):
            lim = draw(st.sampled_from(LIMITS))
            containers.append({"resources": {"limits": ({"aliyun.com/gpu-mem": lim} if lim is not None else {})}})
        pods.append({"metadata": md, "spec": {"nodeName": draw(st.sampled_from(names + ["elsewhere"])), "containers": containers},
                     "status": {"phase": draw(st.sampled_from(["Running", "Pending", "Succeeded", "Failed"]))}})
    return nodes, pods


@settings(max_examples=400, deadline=None)
@given(clusters(), st.booleans(), st.booleans())
def test_product_and_oracle_print_the_same_text(cluster, details, single):
    nodes, pods = cluster
    name = nodes[0]["metadata"]["name"] if single else ""
    assert product(nodes, pods, name, details) == io.inspect(nodes, pods, name, details)


def test_oracle_reproduces_the_frozen_fixture():
    """tests/golden/inspect_cases.json was frozen from the product; the oracle, given the same clusters, must print the
    same text (the fixture's clusters are rebuilt by its generator's own helpers)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mig", os.path.join(ROOT, "tests", "golden", "make_inspect_golden.py"))
    mig = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mig)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "inspect_cases.json")))["cases"]
    from gpushare_device_plugin_b200.testing.mock_kube import make_node
    builders = {"config 1": (179, 8, mig.config1_pods), "pending pseudo-device": (179, 8, mig.awkward_pods), "small GPUs": (16, 2, mig.small_unit_pods)}
    for case in gold:
        slices, n_gpus, pods_fn = next(v for k, v in builders.items() if case["name"].startswith(k))
        n = make_node(mig.NODE, gpu_count=n_gpus)
        n["status"]["allocatable"]["aliyun.com/gpu-mem"] = str(n_gpus * slices)
        n["status"]["addresses"] = [{"type": "Hostname", "address": mig.NODE}, {"type": "InternalIP", "address": "10.0.0.7"}]
        nodes = [n, make_node("cpu-only")]
        argv = case["argv"]
        name = next((a for a in argv if a != "-d"), "")
        assert io.inspect(nodes, pods_fn(), name, "-d" in argv) == case["output"], (case["name"], argv)
