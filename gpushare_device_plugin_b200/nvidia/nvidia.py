"""Inventory + health watch — mirror of pkg/gpu/nvidia/nvidia.go over the C ABI.

Same functions, same argument meaning, same results as the reference; every value comes from
libgpushare_b200.so (one NVML memory query + identity cross-check per GPU instead of go-nvml's
11-getter NewDevice, and an event queue fed by the XID thread and the HBM prober instead of a
per-fake-device NVML registration). Nothing here computes inventory in Python.
"""
from __future__ import annotations

import logging
import threading
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

from .. import device
from .._abi import GSB_EVENT_INVENTORY, GSB_EVENT_PROBE, GSB_EVENT_XID, GSB_PROBE_RECOVERED, GsbError, lib
from . import const

log = logging.getLogger("gpushare.nvidia")

# process globals, as in nvidia.go:15-18 (gpuMemory is set once, from the first GPU, and never reset)
gpuMemory: int = 0
metric: str = const.GiBPrefix


@dataclass
class Device:
    """pluginapi.Device (v1beta1/api.proto:82-90)."""
    ID: str
    Health: str = const.Healthy


def check(err: Exception) -> None:
    """nvidia.go:20-24: any inventory error is fatal (log.Fatalln -> exit status 1)."""
    if err is not None:
        log.critical("Fatal: %s", err)
        raise SystemExit(1)


def generateFakeDeviceID(realID: str, fakeCounter: int) -> str:  # nvidia.go:26-28
    return device.fake_device_id(realID, fakeCounter)


def extractRealDeviceID(fakeDeviceID: str) -> str:  # nvidia.go:30-32
    return device.real_device_id(fakeDeviceID)


def setGPUMemory(raw: int) -> None:  # nvidia.go:34-41
    global gpuMemory
    gpuMemory = device.slices(raw, metric == const.GiBPrefix)
    log.info("set gpu memory: %d", gpuMemory)


def getGPUMemory() -> int:  # nvidia.go:43-45
    return gpuMemory


def getDeviceCount() -> int:  # nvidia.go:47-51
    try:
        return device.device_count()
    except GsbError as e:
        check(e)


def getDevices() -> Tuple[List[Device], Dict[str, int]]:
    """nvidia.go:53-89: fake devices in (GPU index asc, slice asc) order + UUID -> /dev/nvidia minor."""
    try:
        n = device.device_count()
        devs: List[Device] = []
        realDevNames: Dict[str, int] = {}
        for i in range(n):
            d = device.device_info(i)
            log.info("Deivce %s's Path is /dev/nvidia%d", d.uuid, d.minor)
            realDevNames[d.uuid] = d.minor
            log.info("# device Memory: %d", d.total_mib)
            if getGPUMemory() == 0:
                setGPUMemory(d.total_mib)
            for j in range(getGPUMemory()):
                devs.append(Device(ID=device.fake_device_id(d.uuid, j)))
        return devs, realDevNames
    except GsbError as e:
        check(e)


def deviceExists(devs: List[Device], id_: str) -> bool:  # nvidia.go:91-98
    return any(d.ID == id_ for d in devs)


def watchXIDs(stop: threading.Event, devs: List[Device], xids: Callable[[Device], None],
              probe_period_ms: int = 0, window_bytes: int = device.GiB,
              recovered: Optional[Callable[[Device], None]] = None, recovery_cycles: int = 0) -> None:
    """nvidia.go:100-152 with the event source replaced: gsb_health_wait delivers both XID critical
    errors (same NVML event type, registered once per GPU) and verdicts of the active HBM probe.
    The loop shape, the 5 s wait, the XID 31/43/45 filter, "empty UUID = every device" and "every
    fake device of that UUID" are the reference's."""
    lib.gsb_health_set_recovery(recovery_cycles if recovered is not None else 0)
    device.health_start(probe_period_ms, window_bytes)
    try:
        while not stop.is_set():
            e = device.health_wait(5000)  # nvidia.go:126
            if e is None:
                continue
            if e.etype == GSB_EVENT_INVENTORY:  # the low-rate NVML refresh disagrees with what is advertised
                log.warning("inventory of %s changed under the plugin (%d): marking it unhealthy", e.uuid.decode(), e.edata)
            elif e.etype not in (GSB_EVENT_XID, GSB_EVENT_PROBE):  # nvidia.go:127-129
                continue
            if e.etype == GSB_EVENT_XID and lib.gsb_xid_is_benign(e.edata):  # nvidia.go:134-136
                continue
            uuid = e.uuid.decode()
            if e.etype == GSB_EVENT_PROBE and e.edata == GSB_PROBE_RECOVERED:
                # not in the reference (server.go:180 FIXME): only when --health-recovery-cycles > 0
                if recovered is not None and recovery_cycles > 0:
                    for d in devs:
                        if extractRealDeviceID(d.ID) == uuid:
                            recovered(d)
                continue
            if len(uuid) == 0:  # nvidia.go:138-144: all devices are unhealthy
                for d in devs:
                    xids(d)
                continue
            for d in devs:  # nvidia.go:146-150
                if extractRealDeviceID(d.ID) == uuid:
                    xids(d)
    finally:
        device.health_stop()
