"""Kernel sweep on one B200: variant x op x window -> CUDA-event GB/s. Writes gpurun_out/sweep.json.
Diagnostics for DESIGN.md / profiles/; bench.py is the contract."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpushare_device_plugin_b200 import _abi, device  # noqa: E402

GiB, MiB = 1 << 30, 1 << 20
PEAK = 6574.8


def main():
    out = {"rows": []}
    device.init()
    t = time.time()
    arena = device.arena_create(0)
    out["arena_bytes"] = arena
    out["arena_create_s"] = round(time.time() - t, 3)
    print(f"arena {arena} B ({arena / GiB:.2f} GiB) in {out['arena_create_s']} s", flush=True)
    ops = {"fill": _abi.GSB_OP_FILL, "verify": _abi.GSB_OP_VERIFY, "refill": _abi.GSB_OP_VERIFY_REFILL}
    grids = [0] if len(sys.argv) < 2 else [int(g) for g in sys.argv[1].split(",")]
    for wname, wbytes in (("64MiB", 64 * MiB), ("1GiB", GiB), ("16GiB", 16 * GiB), ("full", arena)):
        for vname, variant in (("direct", 1), ("cpasync", 2), ("bulk", 3)):
            for grid in grids:
                seed = 100
                device.probe(0, _abi.GSB_OP_FILL, variant=1, nbytes=wbytes, seed_write=seed)
                for oname, op in ops.items():
                    reps = 3 if wbytes >= 16 * GiB else 10
                    times = []
                    for i in range(reps + 2):
                        # rotate the window across the arena so that nothing is L2-resident
                        n_win = max(1, arena // wbytes)
                        off = ((i * 7) % n_win) * wbytes if wbytes < arena else 0
                        if op != _abi.GSB_OP_FILL and off != 0:
                            device.probe(0, _abi.GSB_OP_FILL, variant=1, offset=off, nbytes=wbytes, seed_write=seed)
                        r = device.probe(0, op, variant=variant, offset=off, nbytes=wbytes, seed_expect=seed,
                                         seed_write=seed, grid=grid)
                        assert r.mismatch_words == 0, (wname, vname, oname)
                        if i >= 2:
                            times.append(r.kernel_ns)
                    best, med = min(times), sorted(times)[len(times) // 2]
                    traffic = wbytes * (2 if op == _abi.GSB_OP_VERIFY_REFILL else 1)
                    row = {"window": wname, "variant": vname, "op": oname, "grid": r.grid_ctas,
                           "best_us": best / 1e3, "median_us": med / 1e3,
                           "gbps_median": traffic / med, "frac_of_measured_peak": traffic / med / PEAK}
                    out["rows"].append(row)
                    print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sweep.json", "w") as f:
        json.dump(out, f, indent=1)
    device.shutdown()


if __name__ == "__main__":
    main()
