"""Thin Python face of the C ABI's device layer (inventory, arena, probe, cycle, health events).

Everything here is a direct call into ``libgpushare_b200.so``; nothing is computed or cached in Python (the
library keeps NVML's (re)start-time inventory snapshot; see include/gpushare_b200.h). Names follow the header.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

from . import _abi
from ._abi import (CycleResult, DeviceInfo, Event, GsbError, ProbeCfg, ProbeResult, check, lib)

GiB = 1 << 30
MiB = 1 << 20


def init() -> None:
    """nvml.Init equivalent (gpumanager.go:36). Raises GsbError; never degrades to a CPU path."""
    check(lib.gsb_init(), "gsb_init")


def shutdown() -> None:
    check(lib.gsb_shutdown(), "gsb_shutdown")


def device_count() -> int:
    n = C.c_uint32(0)
    check(lib.gsb_device_count(C.byref(n)), "gsb_device_count")
    return n.value


@dataclass
class Info:
    index: int
    uuid: str
    bus_id: str
    minor: int
    cuda_ordinal: int
    sm_count: int
    cc: tuple
    total_bytes: int
    total_mib: int
    free_bytes: int
    cuda_total_bytes: int

    @staticmethod
    def of(s: DeviceInfo) -> "Info":
        return Info(s.index, s.uuid.decode(), s.bus_id.decode(), s.minor, s.cuda_ordinal, s.sm_count,
                    (s.cc_major, s.cc_minor), s.total_bytes, s.total_mib, s.free_bytes, s.cuda_total_bytes)


def device_info(idx: int) -> Info:
    s = DeviceInfo()
    check(lib.gsb_device_info_get(idx, C.byref(s)), f"gsb_device_info_get({idx})")
    return Info.of(s)


def inventory_refresh(idx: int = _abi.GSB_ALL_DEVICES) -> None:
    """Live NVML query -> snapshot; what NewNvidiaDevicePlugin's getDevices does at every (re)start (server.go:39)."""
    check(lib.gsb_inventory_refresh(idx), f"gsb_inventory_refresh({idx})")


def inventory_snapshot(idx: int):
    """(Info, age_ns): NVML's last answer as the cycle serves it, identity re-validated on the CUDA side."""
    s, age = DeviceInfo(), C.c_uint64(0)
    check(lib.gsb_inventory_snapshot(idx, C.byref(s), C.byref(age)), f"gsb_inventory_snapshot({idx})")
    return Info.of(s), age.value


def set_option(key: int, value: int) -> None:
    check(lib.gsb_set_option(key, value), f"gsb_set_option({key})")


def get_option(key: int) -> int:
    v = C.c_uint64(0)
    check(lib.gsb_get_option(key, C.byref(v)), f"gsb_get_option({key})")
    return v.value


def test_stall(idx: int, ms: int) -> None:
    """Test hook: keep the device's probe stream busy for `ms` ms (a stand-in for a wedged GPU)."""
    check(lib.gsb_test_stall(idx, ms), f"gsb_test_stall({idx})")


test_stall.__test__ = False  # not a pytest case


def slices(total_mib: int, unit_gib: bool = True) -> int:
    return lib.gsb_slices(total_mib, 1 if unit_gib else 0)


def fake_device_id(uuid: str, j: int) -> str:
    buf = C.create_string_buffer(128)
    n = check(lib.gsb_fake_device_id(uuid.encode(), j, buf, len(buf)), "gsb_fake_device_id")
    return buf.raw[:n].decode()


def real_device_id(fake_id: str) -> str:
    buf = C.create_string_buffer(max(128, len(fake_id) + 1))
    n = check(lib.gsb_real_device_id(fake_id.encode(), buf, len(buf)), "gsb_real_device_id")
    return buf.raw[:n].decode()


def encode_list_and_watch(uuids: Sequence[str], n_slices: int, unhealthy_bits: Optional[bytes] = None) -> bytes:
    arr = (C.c_char_p * len(uuids))(*[u.encode() for u in uuids])
    bits = C.cast(C.c_char_p(unhealthy_bits), C.c_void_p) if unhealthy_bits is not None else None
    need = lib.gsb_encode_list_and_watch(arr, len(uuids), n_slices, bits, None, 0)
    check(int(need), "gsb_encode_list_and_watch(size)")
    buf = C.create_string_buffer(int(need) or 1)
    n = lib.gsb_encode_list_and_watch(arr, len(uuids), n_slices, bits, buf, int(need))
    check(int(n), "gsb_encode_list_and_watch")
    return buf.raw[: int(n)]


def encode_register_request(version: str, endpoint: str, resource_name: str) -> bytes:
    buf = C.create_string_buffer(512)
    n = lib.gsb_encode_register_request(version.encode(), endpoint.encode(), resource_name.encode(), buf, len(buf))
    check(int(n), "gsb_encode_register_request")
    return buf.raw[: int(n)]


def arena_create(idx: int, max_bytes: int = 0, keep_free_bytes: int = 0) -> int:
    out = C.c_uint64(0)
    check(lib.gsb_arena_create(idx, max_bytes, keep_free_bytes, C.byref(out)), f"gsb_arena_create({idx})")
    return out.value


def arena_destroy(idx: int) -> None:
    check(lib.gsb_arena_destroy(idx), f"gsb_arena_destroy({idx})")


def arena_bytes(idx: int) -> int:
    out = C.c_uint64(0)
    check(lib.gsb_arena_bytes(idx, C.byref(out)), f"gsb_arena_bytes({idx})")
    return out.value


def probe(idx: int, op: int, *, variant: int = _abi.GSB_VARIANT_AUTO, offset: int = 0, nbytes: int = 0,
          seed_expect: int = 0, seed_write: int = 0, grid: int = 0, flags: int = _abi.GSB_PROBE_TIMED,
          raise_on_error: bool = True) -> ProbeResult:
    cfg = ProbeCfg(op, variant, offset, nbytes, seed_expect, seed_write, grid, flags)
    res = ProbeResult()
    rc = lib.gsb_probe(idx, C.byref(cfg), C.byref(res))
    if raise_on_error:
        check(rc, f"gsb_probe({idx})")
    return res


def probe_all(idxs: Sequence[int], op: int, **kw) -> List[ProbeResult]:
    cfg = ProbeCfg(op, kw.get("variant", 0), kw.get("offset", 0), kw.get("nbytes", 0), kw.get("seed_expect", 0),
                   kw.get("seed_write", 0), kw.get("grid", 0), kw.get("flags", _abi.GSB_PROBE_TIMED))
    arr = (C.c_uint32 * len(idxs))(*idxs)
    res = (ProbeResult * len(idxs))()
    check(lib.gsb_probe_all(len(idxs), arr, C.byref(cfg), res), "gsb_probe_all")
    return list(res)


def arena_read(idx: int, offset: int, nbytes: int) -> bytes:
    buf = C.create_string_buffer(nbytes)
    check(lib.gsb_arena_read(idx, offset, buf, nbytes), "gsb_arena_read")
    return buf.raw


def arena_write(idx: int, offset: int, data: bytes) -> None:
    check(lib.gsb_arena_write(idx, offset, data, len(data)), "gsb_arena_write")


class Cycler:
    """Reusable buffers around gsb_cycle: one inventory + health-probe cycle of one device."""

    def __init__(self, idx: int, window_bytes: int = GiB, unit_gib: bool = True,
                 variant: int = _abi.GSB_VARIANT_AUTO, lw_cap: int = 1 << 16):
        self.idx, self.window_bytes, self.unit_gib, self.variant = idx, window_bytes, unit_gib, variant
        self.buf = C.create_string_buffer(lw_cap)
        self.res = CycleResult()
        self.cycle_no = 0

    def step(self, raise_on_error: bool = True) -> CycleResult:
        rc = lib.gsb_cycle(self.idx, self.cycle_no, self.window_bytes, 1 if self.unit_gib else 0, self.variant,
                           self.buf, len(self.buf), C.byref(self.res))
        self.rc = rc
        if raise_on_error:
            check(rc, f"gsb_cycle({self.idx})")
        self.cycle_no += 1
        return self.res

    def list_and_watch_bytes(self) -> bytes:
        return self.buf.raw[: self.res.lw_len]


class NodeCycler:
    """gsb_cycle_all: every listed device's cycle concurrently (one persistent native thread per device)
    and the concatenated ListAndWatchResponse of the node."""

    def __init__(self, idxs: Sequence[int], window_bytes: int = GiB, unit_gib: bool = True,
                 variant: int = _abi.GSB_VARIANT_AUTO, lw_cap: int = 1 << 20):
        self.idxs = (C.c_uint32 * len(idxs))(*idxs)
        self.n, self.window_bytes, self.unit_gib, self.variant = len(idxs), window_bytes, unit_gib, variant
        self.buf = C.create_string_buffer(lw_cap)
        self.res = (CycleResult * len(idxs))()
        self.cycle_no, self.lw_len = 0, 0

    def step(self):
        n = lib.gsb_cycle_all(self.n, self.idxs, self.cycle_no, self.window_bytes, 1 if self.unit_gib else 0,
                              self.variant, self.buf, len(self.buf), self.res)
        check(int(n), "gsb_cycle_all")
        self.lw_len = int(n)
        self.cycle_no += 1
        return self.res

    def list_and_watch_bytes(self) -> bytes:
        return self.buf.raw[: self.lw_len]


def health_start(probe_period_ms: int = 0, window_bytes: int = GiB) -> None:
    check(lib.gsb_health_start(probe_period_ms, window_bytes), "gsb_health_start")


def health_stop() -> None:
    check(lib.gsb_health_stop(), "gsb_health_stop")


def health_wait(timeout_ms: int) -> Optional[Event]:
    """≙ nvml.WaitForEvent(set, timeout): an Event, or None on timeout."""
    ev = Event()
    rc = lib.gsb_health_wait(timeout_ms, C.byref(ev))
    if rc in (_abi.GSB_ERR_TIMEOUT, _abi.GSB_ERR_STOPPED):
        return None
    check(rc, "gsb_health_wait")
    return ev


def health_stats(idx: int) -> "_abi.HealthStats":
    """What the prober thread of device `idx` has done since health_start: cycles, sweeps, skips, faults, last sizes."""
    st = _abi.HealthStats()
    check(lib.gsb_health_stats_get(idx, C.byref(st)), f"gsb_health_stats_get({idx})")
    return st


def health_inject(uuid: str, etype: int, edata: int) -> None:
    ev = Event(uuid.encode(), etype, edata)
    check(lib.gsb_health_inject(C.byref(ev)), "gsb_health_inject")
