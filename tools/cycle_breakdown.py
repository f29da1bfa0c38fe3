"""Where does a steady-state cycle's wall time go? Python call wall vs the library's own wall vs
CUDA-event kernel time vs the (overlapped) inventory time; plus raw launch+sync of an EMPTY window as
the floor of the launch path. Diagnostics."""
import ctypes as C
import os
import statistics as st
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpushare_device_plugin_b200 import _abi, device  # noqa: E402

GiB = 1 << 30
device.init()
device.arena_create(0)
cyc = device.Cycler(0, window_bytes=GiB)
for _ in range(20):
    cyc.step()
py, libw, kern, inv = [], [], [], []
for _ in range(500):
    t0 = time.perf_counter_ns()
    r = cyc.step()
    py.append((time.perf_counter_ns() - t0) / 1e3)
    libw.append(r.probe.wall_ns / 1e3)
    kern.append(r.probe.kernel_ns / 1e3)
    inv.append(r.inventory_ns / 1e3)
med = st.median
print(f"python call {med(py):.1f} us | lib probe wall {med(libw):.1f} | kernel (events) {med(kern):.1f} | inventory {med(inv):.1f}")
# floor of the launch path: tiny windows (16 KiB), timed and untimed
for flags, name in ((_abi.GSB_PROBE_TIMED, "timed"), (0, "untimed")):
    ts = []
    for i in range(300):
        t0 = time.perf_counter_ns()
        device.probe(0, _abi.GSB_OP_VERIFY, offset=0, nbytes=16384, seed_expect=0, flags=flags, raise_on_error=False)
        ts.append((time.perf_counter_ns() - t0) / 1e3)
    print(f"16 KiB probe, {name}: python wall median {med(ts):.1f} us, min {min(ts):.1f}")
info_t = []
for _ in range(300):
    t0 = time.perf_counter_ns()
    device.device_info(0)
    info_t.append((time.perf_counter_ns() - t0) / 1e3)
print(f"gsb_device_info_get idle: median {med(info_t):.1f} us")
device.shutdown()
