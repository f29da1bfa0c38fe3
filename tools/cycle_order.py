"""Does the host inventory overlap the kernel? Runs the steady-state cycle under each GSB_CYCLE_ORDER,
quiet and with a deliberately noisy NVML neighbour (nvidia-smi -lms 20), in child processes."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time, statistics as st
sys.path.insert(0, %r)
from gpushare_device_plugin_b200 import device
device.init(); device.arena_create(0, max_bytes=8 << 30)
cyc = device.Cycler(0, window_bytes=1 << 30)
for _ in range(20): cyc.step()
py, kern, inv = [], [], []
for _ in range(400):
    t0 = time.perf_counter_ns(); r = cyc.step(); py.append((time.perf_counter_ns() - t0) / 1e3)
    kern.append(r.probe.kernel_ns / 1e3); inv.append(r.inventory_ns / 1e3)
print(round(st.median(py), 1), round(st.median(kern), 1), round(st.median(inv), 1), round(st.mean(py), 1))
''' % ROOT
names = {0: "launch-then-inventory", 1: "inventory-then-launch", 2: "no-inventory", 3: "170us-host-spin"}
HAMMER = "import pynvml\npynvml.nvmlInit()\nh=pynvml.nvmlDeviceGetHandleByIndex(0)\nwhile True:\n    pynvml.nvmlDeviceGetMemoryInfo(h); pynvml.nvmlDeviceGetPowerUsage(h); pynvml.nvmlDeviceGetClockInfo(h,0)\n"
for noisy in (0, 4, 16):
    bgs = [subprocess.Popen([sys.executable, "-c", HAMMER]) for _ in range(noisy)]
    time.sleep(1.0)
    for order in (0, 1, 2, 3):
        o = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, GSB_CYCLE_ORDER=str(order)),
                           capture_output=True, text=True, timeout=300)
        print(f"hammers={noisy} {names[order]:24s} python-call-median/kernel/inventory/mean us:", o.stdout.strip() or o.stderr[-300:], flush=True)
    for bg in bgs:
        bg.terminate(); bg.wait()
