"""Round-2 tile-shape sweep of the TMA probe kernels (lab build: ./build.sh lab -> libgpushare_b200_lab.so).
One subprocess per (variant, shape): the knobs are read once per process. -> gpurun_out/sweep_r02.json
Questions: (1) does a 64 KiB tile / 1 CTA per SM beat the shipped 32 KiB x 3 / 2 CTAs per SM? (2) which shape should
VERIFY (the transient window's and the start-up walk's read pass) ship with — BULKD 32 KiB x 3 ran it at 0.93?"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LAB = os.path.join(ROOT, "gpushare_device_plugin_b200", "libgpushare_b200_lab.so")
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
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
        for i in range(10 if wb <= 4 * GiB else 5):
            nwin = max(1, arena // wb)
            off = ((i * 5) %% nwin) * wb
            r = device.probe(0, op, variant=variant, offset=off, nbytes=wb, seed_expect=9, seed_write=9)
            assert r.mismatch_words == 0
            if i >= 2: ts.append(r.kernel_ns)
        med = sorted(ts)[len(ts) // 2]
        traffic = wb * (2 if op == 3 else 1)
        rows.append({"bytes": wb, "op": oname, "grid": r.grid_ctas, "median_us": med / 1e3, "gbps": traffic / med, "frac": traffic / med / 6574.8})
print(json.dumps(rows))
''' % ROOT
SHAPES = {0: "16KiBx4", 1: "32KiBx3 (shipped)", 5: "32KiBx6", 6: "64KiBx3", 7: "64KiBx2", 8: "32KiBx4"}
GiB = 1 << 30


def run(env, variant, sizes):
    e = dict(os.environ, GSB_LIB_PATH=LAB)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", CHILD, str(variant), ",".join(map(str, sizes))], env=e, capture_output=True, text=True, timeout=600)
    if out.returncode:
        return {"error": out.stderr[-500:]}
    return json.loads(out.stdout.strip().splitlines()[-1])


def main():
    if not os.path.exists(LAB):
        sys.exit("build the lab library first: ./build.sh lab")
    res = {}

    def show(name, rows):
        res[name] = rows
        if isinstance(rows, dict):
            print(name, rows, flush=True)
        else:
            print(name, " | ".join(f"{r['bytes'] >> 20}MiB {r['op']} {r['median_us']:.0f}us {r['frac']:.3f}" for r in rows), flush=True)
    sizes = [64 << 20, 256 << 20, GiB, 0]
    show("direct", run({}, 1, sizes))
    if os.environ.get("SWEEP_ONLY") == "direct":
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep_r02_direct.json"), "w"), indent=1)
        return
    for vname, v in (("bulk_static", 3), ("bulk_dynamic", 5)):
        for cfg, shape in SHAPES.items():
            show(f"{vname} {shape}", run({"GSB_BULK_CFG": str(cfg), "GSB_DYN_FILL": "1"}, v, sizes))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep_r02.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
