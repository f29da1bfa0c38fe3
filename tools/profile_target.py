"""Small, short target for `ncu --set full` (ncu replays each launch ~40x and saves/restores device
memory around every pass, which it cannot do for the 178 GiB arena): a 2 GiB arena, 1 GiB windows.
usage: profile_target.py <variant 0..5> [n_launches] [refill|transient]
  refill (default): VERIFY_REFILL of 1 GiB windows; transient: the daemon's default pair, FILL then VERIFY of 1 GiB"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpushare_device_plugin_b200 import _abi, device  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mode = sys.argv[3] if len(sys.argv) > 3 else "refill"
device.init()
if mode == "transient":
    device.arena_create(0, max_bytes=1 << 30)
    for i in range(n):
        device.probe(0, _abi.GSB_OP_FILL, variant=variant, seed_write=10 + i)
        r = device.probe(0, _abi.GSB_OP_VERIFY, variant=variant, seed_expect=10 + i)
        assert r.mismatch_words == 0
    print("done", r.kernel_ns)
    device.shutdown()
    sys.exit(0)
device.arena_create(0, max_bytes=2 << 30)
device.probe(0, _abi.GSB_OP_FILL, variant=1, seed_write=1)
seed = 1
for i in range(n):
    r = device.probe(0, _abi.GSB_OP_VERIFY_REFILL, variant=variant, offset=(i % 2) << 30, nbytes=1 << 30,
                     seed_expect=seed if i < 2 else seed - 1, seed_write=seed + 1 if i < 2 else seed)
    if i % 2 == 1:
        seed += 1
print("done", r.kernel_ns)
device.shutdown()
