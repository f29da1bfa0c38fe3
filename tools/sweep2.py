"""Knob sweep (subprocess per setting: the knobs are read once per process). -> gpurun_out/sweep2.json"""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
CHILD = r'''
import json, os, sys
sys.path.insert(0, os.path.dirname(%r))
from gpushare_device_plugin_b200 import _abi, device
GiB = 1 << 30
variant, sizes = int(sys.argv[1]), [int(x) for x in sys.argv[2].split(",")]
device.init()
arena = device.arena_create(0)
rows = []
for w in sizes:
    wb = w or arena
    for oname, op in (("fill", 1), ("verify", 2), ("refill", 3)):
        device.probe(0, 1, variant=3, seed_write=9)
        ts = []
        for i in range(8 if wb <= 4 * GiB else 4):
            nwin = max(1, arena // wb)
            off = ((i * 5) %% nwin) * wb
            r = device.probe(0, op, variant=variant, offset=off, nbytes=wb, seed_expect=9, seed_write=9)
            assert r.mismatch_words == 0
            if i >= 2: ts.append(r.kernel_ns)
        med = sorted(ts)[len(ts) // 2]
        traffic = wb * (2 if op == 3 else 1)
        rows.append({"bytes": wb, "op": oname, "grid": r.grid_ctas, "median_us": med / 1e3, "gbps": traffic / med, "frac": traffic / med / 6574.8})
print(json.dumps(rows))
''' % HERE
def run(env, variant, sizes):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "-c", CHILD, str(variant), ",".join(map(str, sizes))], env=e, capture_output=True, text=True, timeout=600)
    if out.returncode: return {"error": out.stderr[-500:]}
    return json.loads(out.stdout.strip().splitlines()[-1])
res = {}
GiB = 1 << 30
def show(name, rows):
    res[name] = rows
    if isinstance(rows, dict):
        print(name, rows, flush=True)
        return
    print(name, " | ".join(f"{r['bytes'] >> 20}MiB {r['op']} {r['median_us']:.0f}us {r['frac']:.3f}" for r in rows), flush=True)
sizes = [GiB, 16 * GiB, 0]
show("bulk_static", run({}, 3, sizes))
show("bulk_dynamic_prefetched_claim", run({}, 5, sizes))
show("bulk_dynamic_fill_32k6", run({"GSB_DYN_FILL": "1"}, 5, sizes))
show("bulk_dynamic_all_32k3", run({"GSB_DYN_FILL": "1", "GSB_BULK_CFG": "1"}, 5, sizes))
show("bulk_dynamic_all_16k4", run({"GSB_DYN_FILL": "1", "GSB_BULK_CFG": "0"}, 5, sizes))
show("bulk_dynamic_all_16k6", run({"GSB_DYN_FILL": "1", "GSB_BULK_CFG": "2"}, 5, sizes))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/sweep6.json", "w"), indent=1)
