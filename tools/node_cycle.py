"""Config 3 of BASELINE.json in ONE process: all N GPUs of the box inventoried + probed concurrently
by gsb_cycle_all (one persistent native thread, primary context and non-blocking stream per device),
next to the reference's sequential NVML path on the same box. Prints one JSON line.
usage: node_cycle.py [steps] [window_gib (0 = full walk)]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpushare_device_plugin_b200 import device  # noqa: E402

GiB = 1 << 30
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
wgib = int(sys.argv[2]) if len(sys.argv) > 2 else 1
device.init()
n = device.device_count()
arenas = [device.arena_create(i) for i in range(n)]
node = device.NodeCycler(list(range(n)), window_bytes=wgib * GiB)
for _ in range(5):
    node.step()
t0 = time.perf_counter()
kern = [0] * n
inv = [[] for _ in range(n)]  # per device: ns the NVML query + encode took inside the cycle (runs while the kernel does)
per_step = []
for _ in range(steps):
    ts = time.perf_counter_ns()
    res = node.step()
    per_step.append(time.perf_counter_ns() - ts)
    assert all(r.healthy for r in res)
    for i, r in enumerate(res):
        kern[i] += r.probe.kernel_ns
        inv[i].append(r.inventory_ns)
wall = time.perf_counter() - t0
per_step.sort()
w = wgib * GiB if wgib else min(arenas)
out = {"n_gpus": n, "steps": steps, "window_bytes": w, "node_cycles_per_s": steps / wall,
       "device_cycles_per_s": n * steps / wall, "ms_per_node_cycle": wall * 1e3 / steps,
       "kernel_ms_per_cycle_per_device": [k / 1e6 / steps for k in kern],
       "aggregate_hbm_gbs": sum(2 * w * steps / (k / 1e9) / 1e9 for k in kern),
       "lw_bytes": node.lw_len, "devices_advertised": 179 * n,
       # where a slow node cycle spends its time (the round-1 open item: 0.4 ms in some processes, 1-2 ms in others)
       "step_ms": {"min": per_step[0] / 1e6, "p50": per_step[len(per_step) // 2] / 1e6,
                   "p99": per_step[min(len(per_step) - 1, int(len(per_step) * 0.99))] / 1e6, "max": per_step[-1] / 1e6},
       "inventory_us_p50_per_device": [sorted(v)[len(v) // 2] / 1e3 for v in inv],
       "inventory_us_max_per_device": [max(v) / 1e3 for v in inv],
       "knobs": {k: os.environ[k] for k in ("GSB_WORKER_SPIN_US", "GSB_NVML_SERIAL", "GSB_CYCLE_ORDER", "GSB_INVENTORY_POLICY") if k in os.environ},
       "cpus_allowed": len(os.sched_getaffinity(0))}
ref = subprocess.run(["taskset", "-c", "0", os.path.join(ROOT, "oracle", "_ref", "ref_inventory"), "bench", "--iters", "50"],
                     capture_output=True, text=True)
if ref.returncode == 0:
    r = json.loads(ref.stdout.strip().splitlines()[-1])
    out["reference_node_cycles_per_s"] = 1e6 / r["cycle_us"]["mean"]  # inventory + poll; set-up is once per start
    out["reference_phases_us_p50"] = {k: r[k]["p50"] for k in ("inventory_us", "health_setup_us", "health_poll_us")}
    out["reference_register_calls_per_setup"] = r["register_calls_per_setup"]
print(json.dumps(out))
device.shutdown()
