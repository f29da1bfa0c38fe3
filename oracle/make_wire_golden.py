"""Freezes the host-logic oracle: writes tests/golden/allocate_cases.json and tests/golden/inspect_cases.json from
oracle/wire_oracle.py (and cmd/inspect's expected text from oracle-side inputs). TEST INFRASTRUCTURE.

PARITY UNPINNED: the reference has no tests or fixtures on this path and cannot be executed here (Go 1.10, no
toolchain), so these vectors are the ORACLE's answers, not the reference's. What freezing them buys is drift
detection: tests/test_golden_allocate.py holds BOTH the oracle and the product (gsb_allocate through the C ABI)
against the committed file, so a change to either that alters an answer shows up as a diff of this file in review.
Every case names the reference lines whose behaviour it exercises (DESIGN.md has the full oracle <-> Go line map).

Run from the repo root:   python oracle/make_wire_golden.py            (rewrites the files)
                          python oracle/make_wire_golden.py --check    (exit 1 if the files would change)
"""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wire_oracle as wo  # noqa: E402

NODE = "b200-0"
UUIDS = ["GPU-%08x-4820-abfc-e83e-9431819757%02x" % (0xfef80890 + i, i) for i in range(8)]
MINORS8 = dict(zip(UUIDS, [2, 3, 0, 1, 6, 7, 4, 5]))  # minor != index, as on real HGX boxes
T0 = 1_700_000_000_000_000_000


def pod(i, *, node=NODE, limits=(4,), idx="0", assume=T0, assigned="false", uid=None, ns="default", init_limits=None,
        extra_ann=None):
    """One v1.Pod as the apiserver's JSON has it; None for an annotation leaves the key out."""
    ann = {}
    if idx is not None:
        ann[wo.EnvResourceIndex] = str(idx)
    if assume is not None:
        ann[wo.EnvResourceAssumeTime] = str(assume)
    if assigned is not None:
        ann[wo.EnvAssignedFlag] = assigned
    ann.update(extra_ann or {})
    p = {"metadata": {"name": f"pod-{i:02d}", "namespace": ns, "uid": uid or f"uid-{i}", "annotations": ann},
         "spec": {"nodeName": node,
                  "containers": [{"name": f"c{k}", "resources": {"limits": ({wo.resourceName: str(v)} if v is not None else {})}}
                                 for k, v in enumerate(limits)]},
         "status": {"phase": "Pending"}}
    if init_limits:
        p["spec"]["initContainers"] = [{"name": "init", "resources": {"limits": {wo.resourceName: str(v)}}} for v in init_limits]
    return p


def ids(n, g=0):
    return [wo.generateFakeDeviceID(UUIDS[g], j) for j in range(n)]


CASES = []


def case(name, cites, requests, pods, devNameMap=None, slices=179, unit="GiB", cgpu=False):
    devNameMap = MINORS8 if devNameMap is None else devNameMap
    envs, matched = wo.Allocate(requests, copy.deepcopy(pods), NODE, devNameMap, slices, unit, cgpu)
    kind = "matched" if matched is not None else ("single_gpu" if len(devNameMap) == 1 and (
        not envs or not envs[0][wo.envNVGPU].startswith("no-gpu-has")) else "err_response")
    if matched is None and len(devNameMap) == 1 and envs and envs[0][wo.envNVGPU].startswith("no-gpu-has"):
        kind = "err_response"
    CASES.append({"name": name, "reference_lines": cites, "container_requests": requests, "pods": pods,
                  "ctx": {"devNameMap": devNameMap, "slices": slices, "unit": unit, "disable_cgpu_isolation": cgpu, "node": NODE},
                  "want": {"kind": kind, "envs": envs, "matched_uid": matched["metadata"]["uid"] if matched else None,
                           "pod_req_gpu": sum(len(r) for r in requests),
                           "response_hex": wo.marshal_AllocateResponse(envs).hex()}})


# ---- allocate.go:24-39 / 54-56 / 179-184: error response shapes -----------------------------------------------
case("no pods, 8 GPUs -> buildErrResponse", "allocate.go:24-39,179-184", [ids(4)], [])
case("no pods, MiB unit changes the poison text and DEV", "allocate.go:26 (metric), const.go:34-35", [ids(3)], [], slices=183359, unit="MiB")
case("two containers, no candidate: one poison env set per container, CONTAINER differs", "allocate.go:27-37,54-56",
     [ids(2), ids(5, 1)], [])
case("empty request: zero containers, zero responses", "allocate.go:54-56,113", [], [])
case("a container asking for zero devices still gets a response", "allocate.go:113-128", [[], ids(4)],
     [pod(0, limits=(4,), idx="2")])
# ---- allocate.go:78-88: first candidate in assume-time order whose limit == request ----------------------------
case("match: envs for idx 3 (SURVEY 8(c) vector 6)", "allocate.go:113-128", [ids(4)], [pod(0, idx="3", assume=5)])
case("oldest assume time wins, LIST order does not matter", "podmanager.go:241-262, allocate.go:78-88", [ids(4)],
     [pod(0, idx="1", assume=T0 + 9), pod(1, idx="2", assume=T0 + 1), pod(2, idx="3", assume=T0 + 5)])
case("older pod of another size is skipped", "allocate.go:79", [ids(4)],
     [pod(0, limits=(8,), idx="1", assume=T0), pod(1, limits=(4,), idx="2", assume=T0 + 1)])
case("limit summed over ALL containers of the pod", "podutils.go:122-131", [ids(2), ids(4, 1)],
     [pod(0, limits=(2, 4), idx="7", assume=9)])
case("initContainers do not count", "podutils.go:124 (Spec.Containers only)", [ids(4)],
     [pod(0, limits=(4,), idx="1", init_limits=(16,))])
case("a container without the resource contributes 0", "podutils.go:126", [ids(4)], [pod(0, limits=(4, None), idx="6")])
case("cgpu.disable.isolation label -> CGPU_DISABLE=true", "allocate.go:124-126, podmanager.go:59-72", [ids(4)],
     [pod(0, idx="0")], cgpu=True)
# ---- resource.Quantity spellings (podutils.go:127 val.Value()) -------------------------------------------------
for q, n in (("4", 4), ("4000m", 4), ("1k", 1000), ("1Ki", 1024), ("2e1", 20), ("0.5", 1), ("1500m", 2)):
    case(f"quantity {q!r} is Value() == {n}", "podutils.go:127 (resource.Quantity.Value rounds up)", [ids(n)],
         [pod(0, limits=(q,), idx="1")])
# ---- podutils.go:78-119: candidate predicate ---------------------------------------------------------------------
case("no assume-time annotation -> not a candidate", "podutils.go:93-98", [ids(4)], [pod(0, assume=None)])
case("assigned == true -> not a candidate", "podutils.go:101-113", [ids(4)], [pod(0, assigned="true")])
case("assigned flag missing -> not a candidate", "podutils.go:114-118", [ids(4)], [pod(0, assigned=None)])
case("assigned == 'False' (capital) -> not a candidate: exact string compare", "podutils.go:103", [ids(4)], [pod(0, assigned="False")])
case("limit 0 -> not a candidate", "podutils.go:83-88", [ids(4)], [pod(0, limits=(0,))])
case("assume-time annotation present but empty -> candidate with time 0 (sorts first)", "podutils.go:64-75,93",
     [ids(4)], [pod(0, idx="1", assume=T0), pod(1, idx="2", assume="")])
# ---- podutils.go:64-75: ParseUint failures -> 0 -------------------------------------------------------------------
for bad in ("-5", "+5", "12x", "18446744073709551616", " 7", "1e3"):
    case(f"assume time {bad!r} fails ParseUint -> 0 -> oldest", "podutils.go:66-72 (strconv.ParseUint base 10, 64 bit)", [ids(4)],
         [pod(0, idx="1", assume=T0), pod(1, idx="2", assume=bad)])
case("assume time 18446744073709551615 is the largest uint64 and parses", "podutils.go:66", [ids(4)],
     [pod(0, idx="1", assume="18446744073709551615"), pod(1, idx="2", assume=T0)])
# ---- podutils.go:37-61 + allocate.go:91-110: IDX handling ---------------------------------------------------------
for bad, why in ((None, "annotation missing"), ("x7", "Atoi fails"), ("", "Atoi of empty fails"), ("7.0", "Atoi fails"),
                 ("-2", "negative"), ("8", "no GPU with minor 8 on this node"), ("9223372036854775808", "Atoi out of range")):
    case(f"IDX {bad!r}: {why} -> error response, the matched pod is NOT patched", "podutils.go:37-61, allocate.go:91-110", [ids(4)],
         [pod(0, idx=bad), pod(1, idx="1", assume=T0 + 1)])
case("IDX '+3': strconv.Atoi accepts a sign", "podutils.go:45 (strconv.Atoi)", [ids(4)], [pod(0, idx="+3")])
case("IDX '007': leading zeros parse", "podutils.go:45", [ids(4)], [pod(0, idx="007")])
case("IDX is matched against the /dev/nvidia MINOR, not the NVML index", "nvidia.go:65-67, server.go:72-83", [ids(4)],
     [pod(0, idx="6")])
case("pod with no annotations at all is not a candidate", "podutils.go:41,93", [ids(4)],
     [{"metadata": {"name": "bare", "namespace": "default", "uid": "uid-bare"},
       "spec": {"nodeName": NODE, "containers": [{"resources": {"limits": {wo.resourceName: "4"}}}]}, "status": {"phase": "Pending"}}])
# ---- podmanager.go:187-199: node filter and UID dedupe ---------------------------------------------------------------
case("pods of another node are ignored", "podmanager.go:187-193", [ids(4)], [pod(0, node="b200-1", idx="1"), pod(1, idx="2", assume=T0 + 1)])
case("duplicate UID: the first occurrence in LIST order is kept", "podmanager.go:196-199", [ids(4)],
     [pod(0, idx="1", uid="same", assume=T0 + 5), pod(1, idx="2", uid="same", assume=T0)])
case("duplicate UID where the first copy is not a candidate hides the second", "podmanager.go:196-199 then podutils.go:78-119", [ids(4)],
     [pod(0, idx="1", uid="same", assigned="true"), pod(1, idx="2", uid="same")])
# ---- podmanager.go:241-262: order among TIED assume times (Go 1.10 sort.Sort with a non-strict Less) ---------------
for n in (2, 3, 6, 7, 11, 12):
    case(f"{n} candidates with identical assume time: Go 1.10 small-slice path (gap-6 pass + insertion sort, Less is <=)",
         "podmanager.go:256-258; go1.10 sort.go quickSort b-a <= 12", [ids(4)],
         [pod(i, idx=str(i % 8), assume=T0) for i in range(n)])
for n in (13, 20, 41, 64):
    case(f"{n} candidates with identical assume time: Go 1.10 quickSort/doPivot path", "podmanager.go:256-258; go1.10 sort.go doPivot",
         [ids(4)], [pod(i, idx=str(i % 8), assume=T0) for i in range(n)])
case("13 candidates, two distinct times interleaved", "podmanager.go:241-262", [ids(4)],
     [pod(i, idx=str(i % 8), assume=T0 + (i % 2)) for i in range(13)])
case("config 4: 64 pending pods @4 GiB, first call", "SURVEY 8(d) config 4", [ids(4)],
     [pod(i, idx=str(i // 8), assume=T0 + i) for i in range(64)], devNameMap={u: i for i, u in enumerate(UUIDS)})
# ---- allocate.go:151-177: single-GPU shortcut ------------------------------------------------------------------------
ONE = {UUIDS[0]: 6}
case("one GPU, no candidate: NVIDIA_VISIBLE_DEVICES is the UUID, IDX the minor", "allocate.go:151-177", [ids(2)], [], devNameMap=ONE)
case("one GPU, candidate present: the normal path wins (index, not UUID)", "allocate.go:90,151", [ids(2)],
     [pod(0, limits=(2,), idx="6")], devNameMap=ONE)
case("one GPU, candidate with a bad IDX: error response, NOT the shortcut (found == true)", "allocate.go:90-110", [ids(2)],
     [pod(0, limits=(2,), idx="0")], devNameMap=ONE)
case("one GPU, two containers, cgpu label", "allocate.go:160-174", [ids(1), ids(2)], [], devNameMap=ONE, cgpu=True)

# ---- cmd/inspect: the oracle-side inputs of tests/test_inspect.py are frozen by that test's own expected text --------


def main():
    out = {"_comment": "generated by oracle/make_wire_golden.py from oracle/wire_oracle.py. PARITY UNPINNED: these are the "
                       "oracle's answers (a hand-read restatement of allocate.go / podutils.go / podmanager.go), frozen for "
                       "drift detection; no reference-produced vector exists for this path.",
           "node": NODE, "cases": CASES}
    text = json.dumps(out, indent=1, sort_keys=True) + "\n"
    path = os.path.join(ROOT, "tests", "golden", "allocate_cases.json")
    if "--check" in sys.argv:
        same = os.path.exists(path) and open(path).read() == text
        print("allocate_cases.json:", "up to date" if same else "WOULD CHANGE")
        sys.exit(0 if same else 1)
    with open(path, "w") as f:
        f.write(text)
    print(f"wrote {path}: {len(CASES)} cases")


if __name__ == "__main__":
    main()
