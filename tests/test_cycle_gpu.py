"""Full-size properties (BASELINE sizes): whole-arena walk, window rotation, health verdicts."""
import ctypes as C

import numpy as np
import pytest

from oracle import probe_oracle as po
from oracle import wire_oracle as wo

pytestmark = pytest.mark.gpu

from gpushare_device_plugin_b200 import _abi  # noqa: E402

GiB = 1 << 30


@pytest.fixture(scope="module")
def full_arena(gsb):
    nbytes = gsb.arena_create(0)  # everything the driver will hand out
    yield nbytes
    gsb.arena_destroy(0)


def c_checksums(c_oracle, first_word, n_words, seed):
    res = (C.c_uint32 * 2)()
    c_oracle.po_pattern_checksums(C.c_uint64(first_word), C.c_uint64(n_words), C.c_uint32(seed), res)
    return res[0], res[1]


def test_arena_is_what_is_allocatable(gsb, full_arena):
    info = gsb.device_info(0)
    assert full_arena % (2 << 20) == 0
    assert 170 * GiB < full_arena <= info.cuda_total_bytes
    assert gsb.arena_bytes(0) == full_arena
    # with the arena mapped the device has (almost) nothing left
    assert gsb.device_info(0).free_bytes < 2 * GiB


@pytest.mark.parametrize("variant", [_abi.GSB_VARIANT_DIRECT, _abi.GSB_VARIANT_CPASYNC, _abi.GSB_VARIANT_BULK, _abi.GSB_VARIANT_BULKW, _abi.GSB_VARIANT_BULKD],
                         ids=["direct", "cpasync", "bulk", "bulkw", "bulkd"])
def test_full_walk_checksum_is_exact(gsb, full_arena, c_oracle, variant):
    """One launch over ~178 GiB (word indices cross 2^32 at 64 GiB): checksum == the oracle's over the
    whole arena, and == the fold of 7 ragged sub-window launches (linearity)."""
    seed = 1000 + variant
    gsb.probe(0, _abi.GSB_OP_FILL, variant=variant, seed_write=seed)
    r = gsb.probe(0, _abi.GSB_OP_VERIFY, variant=variant, seed_expect=seed)
    assert r.bytes_walked == full_arena and r.mismatch_words == 0
    assert (r.checksum_xor, r.checksum_sum) == c_checksums(c_oracle, 0, full_arena // 16, seed)
    cuts = [0, 16 * 999, 3 * GiB + 48, 64 * GiB - 32, 64 * GiB + 16 * 77, 100 * GiB, 170 * GiB + 16, full_arena]
    parts = []
    for a, b in zip(cuts, cuts[1:]):
        p = gsb.probe(0, _abi.GSB_OP_VERIFY, variant=variant, offset=a, nbytes=b - a, seed_expect=seed)
        assert p.mismatch_words == 0
        parts.append((p.checksum_xor, p.checksum_sum))
    assert po.fold(parts) == (r.checksum_xor, r.checksum_sum)


def test_bytes_around_the_64gib_word_boundary(gsb, full_arena):
    gsb.probe(0, _abi.GSB_OP_FILL, seed_write=77)
    for off in (64 * GiB - 4096, full_arena - 8192, 128 * GiB - 16 * 5):
        got = np.frombuffer(gsb.arena_read(0, off, 8192), dtype=np.uint32).reshape(-1, 4)
        assert np.array_equal(got, po.pattern(off // 16, 512, 77))


def test_single_fault_in_178gib_is_found(gsb, full_arena):
    gsb.probe(0, _abi.GSB_OP_FILL, seed_write=31)
    off = 97 * GiB + 16 * 123457
    cur = np.frombuffer(gsb.arena_read(0, off, 16), dtype=np.uint32).copy()
    cur[2] ^= np.uint32(1 << 19)
    gsb.arena_write(0, off, cur.tobytes())
    r = gsb.probe(0, _abi.GSB_OP_VERIFY_REFILL, seed_expect=31, seed_write=32)
    assert (r.mismatch_words, r.mismatch_bits, r.first_bad_offset) == (1, 1, off)
    assert gsb.probe(0, _abi.GSB_OP_VERIFY, seed_expect=32).mismatch_words == 0


def test_cycle_rotates_windows_and_reports_reference_bytes(gsb, full_arena):
    gsb.arena_create(0)  # fresh generations
    cyc = gsb.Cycler(0, window_bytes=GiB)
    info = gsb.device_info(0)
    want_lw = wo.marshal_ListAndWatchResponse(wo.getDevices(
        [{"uuid": info.uuid, "path": f"/dev/nvidia{info.minor}", "memory_mib": info.total_mib}])[0])
    n_win = gsb.arena_bytes(0) // GiB
    for k in range(n_win + 3):  # wraps: the first windows are re-verified against their refill
        res = cyc.step()
        assert res.healthy == 1 and res.slices == 179
        assert res.probe.bytes_walked == GiB and res.probe.bytes_read == GiB and res.probe.bytes_written == GiB
        assert res.probe.mismatch_words == 0
        assert cyc.list_and_watch_bytes() == want_lw
    # whole-arena cycle over windows that now carry different generations (seed table path)
    whole = gsb.Cycler(0, window_bytes=0)
    res = whole.step()
    assert res.healthy == 1 and res.probe.bytes_walked == gsb.arena_bytes(0)


def test_cycle_flags_corruption_and_unhealthy_is_sticky(gsb, full_arena):
    gsb.arena_create(0)
    cyc = gsb.Cycler(0, window_bytes=GiB)
    assert cyc.step().healthy == 1  # window 0
    gsb.arena_write(0, GiB + 4096, b"\xff" * 16)  # inside window 1
    res = cyc.step()
    assert res.healthy == 0 and res.probe.mismatch_words == 1 and res.probe.first_bad_offset == GiB + 4096
    devs = wo.unmarshal_ListAndWatchResponse(cyc.list_and_watch_bytes())
    assert len(devs) == 179 and all(h == wo.Unhealthy for _, h in devs)
    # FIXME-parity (server.go:180): no way back to Healthy, even though window 2 is clean
    res = cyc.step()
    assert res.probe.mismatch_words == 0
    assert all(h == wo.Unhealthy for _, h in wo.unmarshal_ListAndWatchResponse(cyc.list_and_watch_bytes()))
    gsb.shutdown()  # drop the sticky state for whoever runs next
    gsb.init()


def need_gpus(gsb, k):
    n = gsb.device_count()
    if n < k:
        pytest.skip(f"needs >= {k} GPUs in one process, this box has {n} (run under gpurun --gpus {k})")
    return n


@pytest.mark.parametrize("multi", [False, True], ids=["one-device", "all-devices"])
def test_probe_all_one_thread_per_device(gsb, multi):
    n = need_gpus(gsb, 2) if multi else 1
    for i in range(n):
        gsb.arena_create(i, max_bytes=2 * GiB)
    gsb.probe_all(list(range(n)), _abi.GSB_OP_FILL, seed_write=1)
    rs = gsb.probe_all(list(range(n)), _abi.GSB_OP_VERIFY_REFILL, seed_expect=1, seed_write=2)
    assert len(rs) == n and all(r.status == 0 and r.mismatch_words == 0 and r.bytes_walked == 2 * GiB for r in rs)
    for i in range(n):
        gsb.arena_destroy(i)


@pytest.mark.parametrize("multi", [False, True], ids=["one-device", "all-devices"])
def test_node_cycle_all_devices_in_one_process(gsb, multi):
    """gsb_cycle_all: every GPU of the box probed concurrently (own thread/context/stream), one joined
    ListAndWatchResponse — equal to the reference's sequential getDevices + marshal. The all-devices case needs a
    multi-GPU box and SKIPS on a 1-GPU one (it used to pass there vacuously with n == 1)."""
    n = need_gpus(gsb, 2) if multi else 1
    for i in range(n):
        gsb.arena_create(i, max_bytes=4 * GiB)
    infos = [gsb.device_info(i) for i in range(n)]
    want = wo.marshal_ListAndWatchResponse(wo.getDevices(
        [{"uuid": d.uuid, "path": f"/dev/nvidia{d.minor}", "memory_mib": d.total_mib} for d in infos])[0])
    node = gsb.NodeCycler(list(range(n)), window_bytes=GiB)
    for k in range(6):  # wraps the 4 windows
        res = node.step()
        assert all(r.healthy == 1 and r.probe.bytes_walked == GiB and r.slices == 179 for r in res)
        assert node.list_and_watch_bytes() == want
    # corrupt the LAST device's next window: only its fake devices flip
    victim = n - 1
    gsb.arena_write(victim, (6 % 4) * GiB + 160, b"\x00" * 16)
    res = node.step()
    assert [r.healthy for r in res] == [1] * (n - 1) + [0]
    devs = wo.unmarshal_ListAndWatchResponse(node.list_and_watch_bytes())
    bad = {wo.extractRealDeviceID(i) for i, h in devs if h == wo.Unhealthy}
    assert bad == {infos[victim].uuid} and len(devs) == 179 * n
    # sticky in the node cycle too (server.go:180): the next windows are clean, the GPU stays Unhealthy; and the
    # fake devices of the other GPUs never flip (nvidia.go:146-150)
    res = node.step()
    assert res[victim].probe.mismatch_words == 0 and [r.healthy for r in res] == [1] * (n - 1) + [0]
    devs = wo.unmarshal_ListAndWatchResponse(node.list_and_watch_bytes())
    assert {wo.extractRealDeviceID(i) for i, h in devs if h == wo.Unhealthy} == {infos[victim].uuid}
    # a device listed twice is refused, not raced
    with pytest.raises(_abi.GsbError):
        gsb.NodeCycler([0, 0], window_bytes=GiB).step()
    for i in range(n):
        gsb.arena_destroy(i)
    gsb.shutdown()
    gsb.init()


def test_prober_reports_a_fault_and_optionally_a_recovery(gsb):
    """The daemon's prober thread on a real GPU: a corrupted word in its window raises a PROBE event; with
    recovery enabled, K clean cycles later (the refill repaired the word) a RECOVERED event follows."""
    from gpushare_device_plugin_b200._abi import lib
    gsb.arena_create(0, max_bytes=2 * GiB)
    lib.gsb_health_set_recovery(5)
    gsb.health_start(probe_period_ms=5, window_bytes=GiB)
    try:
        assert gsb.health_wait(300) is None  # clean windows: silence
        gsb.arena_write(0, 4096, b"\x5a" * 16)
        gsb.arena_write(0, GiB + 4096, b"\x5a" * 16)  # whichever window comes next
        ev = gsb.health_wait(5000)
        assert ev is not None and ev.etype == 0x100 and ev.edata == 1 and ev.uuid.decode() == gsb.device_info(0).uuid
        ev2 = gsb.health_wait(5000)
        assert ev2 is not None and (ev2.etype, ev2.edata) == (0x100, 3)
    finally:
        gsb.health_stop()
        lib.gsb_health_set_recovery(0)
        gsb.arena_destroy(0)
        gsb.shutdown()
        gsb.init()
