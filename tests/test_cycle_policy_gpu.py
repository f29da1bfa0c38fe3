"""Round-2 behaviour of the cycle on a real B200, through the C ABI: the inventory snapshot (NVML asked at
(re)start, identity re-validated per cycle on the CUDA side), the transient probe window (allocate -> fill ->
verify -> free: nothing held between cycles), and the completion watchdog (a launch that never finishes must not
keep the GPU Healthy)."""
import time

import pytest

from oracle import wire_oracle as wo

pytestmark = pytest.mark.gpu

from gpushare_device_plugin_b200 import _abi  # noqa: E402

GiB = 1 << 30


def nvml_free(idx=0):
    import pynvml
    pynvml.nvmlInit()
    return pynvml.nvmlDeviceGetMemoryInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).free


def test_snapshot_cycle_serves_the_reference_inventory_without_asking_nvml(gsb):
    gsb.arena_create(0, max_bytes=4 * GiB)
    live = gsb.device_info(0)  # a live NVML query; also rewrites the snapshot
    snap, age = gsb.inventory_snapshot(0)
    assert (snap.uuid, snap.minor, snap.total_bytes, snap.total_mib, snap.bus_id) == \
        (live.uuid, live.minor, live.total_bytes, live.total_mib, live.bus_id) and age < 5e9
    assert gsb.get_option(_abi.GSB_OPT_INVENTORY_POLICY) == _abi.GSB_INVENTORY_SNAPSHOT
    cyc = gsb.Cycler(0, window_bytes=GiB)
    want = wo.marshal_ListAndWatchResponse(wo.getDevices(
        [{"uuid": live.uuid, "path": f"/dev/nvidia{live.minor}", "memory_mib": live.total_mib}])[0])
    inv = []
    for _ in range(20):
        r = cyc.step()
        assert r.healthy == 1 and r.inventory_live == 0 and r.transient == 0 and r.snapshot_age_ns > 0
        assert (r.info.uuid.decode(), r.info.minor, r.info.total_bytes, r.slices) == (live.uuid, live.minor, live.total_bytes, 179)
        assert cyc.list_and_watch_bytes() == want  # byte-identical to the reference's list, from the snapshot
        inv.append(r.inventory_ns)
    assert sorted(inv)[len(inv) // 2] < 100_000  # no driver round trip on the cycle's path (was 6 us .. 2.3 ms)
    # the same cycle with the round-1 policy asks NVML itself and agrees bit for bit
    gsb.set_option(_abi.GSB_OPT_INVENTORY_POLICY, _abi.GSB_INVENTORY_LIVE)
    try:
        r = cyc.step()
        assert r.inventory_live == 1 and r.info.total_bytes == live.total_bytes and cyc.list_and_watch_bytes() == want
    finally:
        gsb.set_option(_abi.GSB_OPT_INVENTORY_POLICY, _abi.GSB_INVENTORY_SNAPSHOT)
    # refresh: the snapshot gets younger, the answer stays
    time.sleep(0.05)
    _, age1 = gsb.inventory_snapshot(0)
    gsb.inventory_refresh()
    snap2, age2 = gsb.inventory_snapshot(0)
    assert age2 < age1 and snap2.total_bytes == live.total_bytes and snap2.free_bytes < live.total_bytes
    gsb.arena_destroy(0)


def test_transient_window_holds_nothing_between_cycles(gsb):
    with pytest.raises(_abi.GsbError):
        gsb.arena_bytes(0)  # no standing arena
    free0 = nvml_free()
    cyc = gsb.Cycler(0, window_bytes=GiB)
    live = gsb.device_info(0)
    want = wo.marshal_ListAndWatchResponse(wo.getDevices(
        [{"uuid": live.uuid, "path": f"/dev/nvidia{live.minor}", "memory_mib": live.total_mib}])[0])
    for _ in range(5):
        r = cyc.step()
        assert r.healthy == 1 and r.transient == 1 and r.slices == 179 and cyc.list_and_watch_bytes() == want
        assert r.probe.status == 0 and r.probe.mismatch_words == 0
        assert (r.probe.bytes_walked, r.probe.bytes_read, r.probe.bytes_written) == (GiB, GiB, GiB)
        assert r.probe.kernel_ns > 0
        with pytest.raises(_abi.GsbError):
            gsb.arena_bytes(0)  # given back
    assert abs(nvml_free() - free0) < 64 << 20  # and NVML agrees: the window is not held


def test_transient_window_with_nothing_allocatable_is_silence_not_a_fault(gsb):
    gsb.set_option(_abi.GSB_OPT_TRANSIENT_KEEP_FREE_BYTES, 1 << 50)  # "tenants hold everything"
    try:
        cyc = gsb.Cycler(0, window_bytes=GiB)
        r = cyc.step()
        assert cyc.rc == 0 and r.healthy == 1 and r.transient == 1 and r.slices == 179
        assert r.probe.status == _abi.GSB_ERR_OUT_OF_MEMORY and r.probe.bytes_walked == 0
        assert len(wo.unmarshal_ListAndWatchResponse(cyc.list_and_watch_bytes())) == 179
    finally:
        gsb.set_option(_abi.GSB_OPT_TRANSIENT_KEEP_FREE_BYTES, GiB)


def test_prober_runs_transient_windows_silently_and_releases_them(gsb):
    free0 = nvml_free()
    gsb.health_start(probe_period_ms=10, window_bytes=GiB)
    try:
        assert gsb.health_wait(1000) is None  # clean transient windows: no event
    finally:
        gsb.health_stop()
    assert abs(nvml_free() - free0) < 64 << 20


def test_watchdog_turns_a_launch_that_never_finishes_into_a_verdict(gsb):
    """A stalled stream stands in for a wedged GPU: the probe queued behind it cannot finish inside the watchdog, so
    the cycle returns GSB_ERR_TIMEOUT with healthy == 0 instead of parking the thread in cudaStreamSynchronize; while
    the stall lasts nothing more is queued; once the stream drains the device probes clean again."""
    assert gsb.get_option(_abi.GSB_OPT_WATCHDOG_MS) == 2000  # on by default
    gsb.arena_create(0, max_bytes=2 * GiB)
    cyc = gsb.Cycler(0, window_bytes=GiB)
    assert cyc.step().healthy == 1
    gsb.set_option(_abi.GSB_OPT_WATCHDOG_MS, 100)
    try:
        gsb.test_stall(0, 1500)
        t0 = time.monotonic()
        r = cyc.step(raise_on_error=False)
        took = time.monotonic() - t0
        assert cyc.rc == _abi.GSB_ERR_TIMEOUT and r.healthy == 0 and r.probe.status == _abi.GSB_ERR_TIMEOUT
        assert 0.09 < took < 0.6, took
        assert all(h == wo.Unhealthy for _, h in wo.unmarshal_ListAndWatchResponse(cyc.list_and_watch_bytes()))
        t0 = time.monotonic()
        r = gsb.probe(0, _abi.GSB_OP_VERIFY, flags=3, raise_on_error=False)  # still stalled: refused at once
        assert r.status == _abi.GSB_ERR_TIMEOUT and time.monotonic() - t0 < 0.05
        time.sleep(1.6)  # the stall (and the probe queued behind it) drain
        r = gsb.probe(0, _abi.GSB_OP_VERIFY, flags=3)
        assert r.status == 0 and r.mismatch_words == 0 and r.bytes_walked == 2 * GiB
    finally:
        gsb.set_option(_abi.GSB_OPT_WATCHDOG_MS, 2000)
        gsb.arena_destroy(0)
        gsb.shutdown()  # drop the sticky Unhealthy
        gsb.init()


def test_prober_reports_a_wedged_device(gsb):
    gsb.arena_create(0, max_bytes=2 * GiB)
    gsb.set_option(_abi.GSB_OPT_WATCHDOG_MS, 100)
    gsb.health_start(probe_period_ms=10, window_bytes=GiB)
    try:
        assert gsb.health_wait(300) is None
        gsb.test_stall(0, 1200)
        ev = gsb.health_wait(3000)
        assert ev is not None and (ev.etype, ev.edata) == (_abi.GSB_EVENT_PROBE, _abi.GSB_PROBE_FAULT_WEDGED)
        assert ev.uuid.decode() == gsb.device_info(0).uuid
        time.sleep(1.3)
    finally:
        gsb.health_stop()
        gsb.set_option(_abi.GSB_OPT_WATCHDOG_MS, 2000)
        gsb.arena_destroy(0)
        gsb.shutdown()
        gsb.init()


def test_offpath_refresher_reports_an_inventory_that_changed_under_the_snapshot(gsb):
    """While health runs, NVML is re-asked off the cycle's path every GSB_OPT_INVENTORY_REFRESH_MS; an answer that
    differs from the snapshot is an event, never silently absorbed. (The snapshot is skewed by a test hook: a real
    box will not change its memory size on request.)"""
    from gpushare_device_plugin_b200._abi import lib
    info = gsb.device_info(0)
    gsb.set_option(_abi.GSB_OPT_INVENTORY_REFRESH_MS, 100)
    gsb.health_start(probe_period_ms=0, window_bytes=0)
    try:
        assert gsb.health_wait(700) is None  # refreshes that agree with the snapshot are silent
        gsb.set_option(_abi.GSB_OPT_INVENTORY_REFRESH_MS, 1500)  # room to look at the skewed snapshot before the next pass
        time.sleep(0.3)
        assert lib.gsb_test_skew_snapshot(0, info.total_bytes - (1 << 30)) == 0
        snap, _ = gsb.inventory_snapshot(0)
        seen_skewed = snap.total_bytes == info.total_bytes - (1 << 30)  # (a refresh may already have run: then skip the look)
        cyc = gsb.Cycler(0, window_bytes=64 << 20)
        if seen_skewed:
            r = cyc.step()
            assert r.slices in (178, 179)  # 178 = the cycle served the skewed snapshot (179 only if a refresh slipped in between)
        ev = gsb.health_wait(5000)       # until the refresher sees NVML disagree with the snapshot
        assert ev is not None and (ev.etype, ev.edata) == (_abi.GSB_EVENT_INVENTORY, _abi.GSB_INVENTORY_TOTAL_CHANGED)
        assert ev.uuid.decode() == info.uuid
        snap, _ = gsb.inventory_snapshot(0)
        assert snap.total_bytes == info.total_bytes and cyc.step().slices == 179  # and the snapshot is NVML's answer again
    finally:
        gsb.health_stop()
        gsb.set_option(_abi.GSB_OPT_INVENTORY_REFRESH_MS, 5000)


def test_transient_window_kept_by_a_wedged_launch_is_given_back_once_the_stream_drains(gsb):
    """The transient window cannot be unmapped under a launch that has not finished: it is kept, the cycle's verdict is
    the wedge, nothing more is allocated or queued while the wedge lasts — and the first cycle after the stream drains
    gives the window back and probes a fresh one (it must not live on as a standing arena)."""
    with pytest.raises(_abi.GsbError):
        gsb.arena_bytes(0)
    free0 = nvml_free()
    gsb.set_option(_abi.GSB_OPT_WATCHDOG_MS, 100)
    cyc = gsb.Cycler(0, window_bytes=GiB)
    try:
        assert cyc.step().healthy == 1
        gsb.test_stall(0, 1200)
        r = cyc.step(raise_on_error=False)  # the window's FILL queues behind the stall and outlives the watchdog
        assert cyc.rc == _abi.GSB_ERR_TIMEOUT and r.healthy == 0 and r.transient == 1
        assert gsb.arena_bytes(0) == GiB    # kept: it cannot be unmapped under the launch
        t0 = time.monotonic()
        r = cyc.step(raise_on_error=False)  # still wedged: refused at once, nothing more taken
        assert cyc.rc == _abi.GSB_ERR_TIMEOUT and time.monotonic() - t0 < 0.05 and gsb.arena_bytes(0) == GiB
        time.sleep(1.5)
        r = cyc.step(raise_on_error=False)  # drained: the old window is released, a fresh one is walked clean
        assert cyc.rc == 0 and r.transient == 1 and r.probe.status == 0 and r.probe.mismatch_words == 0
        assert r.probe.bytes_walked == GiB and r.healthy == 0  # (Unhealthy stays sticky, server.go:180)
        with pytest.raises(_abi.GsbError):
            gsb.arena_bytes(0)
        assert abs(nvml_free() - free0) < 64 << 20
    finally:
        gsb.set_option(_abi.GSB_OPT_WATCHDOG_MS, 2000)
        gsb.shutdown()
        gsb.init()


def test_shutdown_does_not_wait_for_a_wedged_device(gsb):
    """cudaFree / cudaStreamDestroy wait for the device; under a launch that never ends they would hang the daemon's
    shutdown (SIGTERM, SIGHUP restart) for ever. A wedged device's resources are left to the process instead."""
    gsb.arena_create(0, max_bytes=128 << 20)
    gsb.set_option(_abi.GSB_OPT_WATCHDOG_MS, 100)
    cyc = gsb.Cycler(0, window_bytes=64 << 20)
    assert cyc.step().healthy == 1
    gsb.test_stall(0, 2500)
    cyc.step(raise_on_error=False)
    assert cyc.rc == _abi.GSB_ERR_TIMEOUT
    t0 = time.monotonic()
    gsb.shutdown()
    took = time.monotonic() - t0
    gsb.init()
    assert took < 1.0, took  # the stall had ~2.3 s to go
    assert gsb.get_option(_abi.GSB_OPT_WATCHDOG_MS) == 100  # options are process-wide, not per init
    gsb.set_option(_abi.GSB_OPT_WATCHDOG_MS, 2000)
    time.sleep(2.5)  # let the orphaned stall drain before the next test touches the device
    gsb.arena_create(0, max_bytes=64 << 20)
    assert gsb.probe(0, _abi.GSB_OP_VERIFY, flags=3).mismatch_words == 0
    gsb.arena_destroy(0)


def test_periodic_sweep_walks_everything_allocatable_and_gives_it_back(gsb):
    """GSB_OPT_SWEEP_EVERY_CYCLES: with no standing arena, every Nth prober cycle is a transient window as large as
    whatever is allocatable at that moment (minus the keep-free margin): all free HBM is walked — the coverage of the
    start-up walk, at run time — and given back; the cycles in between stay one window."""
    with pytest.raises(_abi.GsbError):
        gsb.arena_bytes(0)
    free0 = nvml_free()
    gsb.set_option(_abi.GSB_OPT_SWEEP_EVERY_CYCLES, 4)
    gsb.health_start(probe_period_ms=10, window_bytes=GiB)
    try:
        deadline = time.monotonic() + 20
        while gsb.health_stats(0).sweeps < 2 and time.monotonic() < deadline:
            time.sleep(0.1)
        st = gsb.health_stats(0)
        assert st.sweeps >= 2 and st.cycles >= 8 and st.faults == 0 and st.skipped == 0
        assert st.last_sweep_bytes > 150 * GiB and st.last_sweep_bytes <= free0 - GiB + (64 << 20)  # all but the margin
        assert st.last_sweep_ns < 5e9
        assert gsb.health_wait(0) is None  # clean sweeps are silent
    finally:
        gsb.health_stop()
        gsb.set_option(_abi.GSB_OPT_SWEEP_EVERY_CYCLES, 0)
    assert abs(nvml_free() - free0) < 64 << 20  # nothing kept
    with pytest.raises(_abi.GsbError):
        gsb.arena_bytes(0)
