"""gsb_allocate alone (no sockets): ns per decision on an n-pod table. Host-only; python tools/allocate_microbench.py"""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
from gpushare_device_plugin_b200._abi import lib  # noqa: E402
from gpushare_device_plugin_b200.nvidia.allocate import AllocateContext, pod_table  # noqa: E402
from gpushare_device_plugin_b200.testing.mock_kube import config4_pods  # noqa: E402
from oracle import wire_oracle as wo  # noqa: E402  (request bytes only)

uuids = {f"GPU-{i:08d}-0000-0000-0000-000000000000": i for i in range(8)}
actx = AllocateContext(uuids, 179, True, False)
req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
buf = C.create_string_buffer(1 << 16)
n, idx, preq = C.c_size_t(0), C.c_int32(-1), C.c_uint32(0)
for n_pods in (12, 64, 1024, 8192):
    pods = config4_pods("b200-0", n_pods, mod=True)
    table, keep = pod_table(pods, "b200-0")
    reps = max(200, 200000 // n_pods)
    for unique in (0, 1):
        actx.ctx.pods_unique = unique
        t0 = time.perf_counter_ns()
        for _ in range(reps):
            k = lib.gsb_allocate(C.byref(actx.ctx), table, n_pods, req, len(req), buf, 1 << 16, C.byref(n), C.byref(idx), C.byref(preq))
        dt = (time.perf_counter_ns() - t0) / reps
        print(f"pods={n_pods:5d} pods_unique={unique}  {dt/1e3:8.2f} us/decision  kind={k} pod_index={idx.value}")
