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
for _ in range(steps):
    res = node.step()
    assert all(r.healthy for r in res)
    for i, r in enumerate(res):
        kern[i] += r.probe.kernel_ns
wall = time.perf_counter() - t0
w = wgib * GiB if wgib else min(arenas)
out = {"n_gpus": n, "steps": steps, "window_bytes": w, "node_cycles_per_s": steps / wall,
       "device_cycles_per_s": n * steps / wall, "ms_per_node_cycle": wall * 1e3 / steps,
       "kernel_ms_per_cycle_per_device": [k / 1e6 / steps for k in kern],
       "aggregate_hbm_gbs": sum(2 * w * steps / (k / 1e9) / 1e9 for k in kern),
       "lw_bytes": node.lw_len, "devices_advertised": 179 * n}
ref = subprocess.run(["taskset", "-c", "0", os.path.join(ROOT, "oracle", "_ref", "ref_inventory"), "bench", "--iters", "50"],
                     capture_output=True, text=True)
if ref.returncode == 0:
    r = json.loads(ref.stdout.strip().splitlines()[-1])
    out["reference_node_cycles_per_s"] = 1e6 / r["cycle_us"]["mean"]
    out["reference_phases_us_p50"] = {k: r[k]["p50"] for k in ("inventory_us", "health_setup_us", "health_poll_us")}
    out["reference_register_calls_per_cycle"] = r["register_calls_per_cycle"]
print(json.dumps(out))
device.shutdown()
