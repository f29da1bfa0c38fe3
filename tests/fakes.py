"""Synthetic inventory for GPU-less tests of the host side: replaces ONLY the two calls that need a
driver (device enumeration, health-thread start); IDs, slice arithmetic, wire bytes and the Allocate
decision still run in the C ABI."""

from gpushare_device_plugin_b200 import device
from gpushare_device_plugin_b200.nvidia import const, nvidia

UUIDS = ["GPU-%08x-4820-abfc-e83e-9431819757%02x" % (0xfef80890 + i, i) for i in range(8)]
MINORS = [2, 3, 0, 1, 6, 7, 4, 5]
TOTAL_MIB = 183359


def install(monkeypatch, n_gpus=8, metric=const.GiBPrefix):
    nvidia.gpuMemory = 0
    nvidia.metric = metric

    def getDevices():
        devs, names = [], {}
        for i in range(n_gpus):
            names[UUIDS[i]] = MINORS[i]
            if nvidia.getGPUMemory() == 0:
                nvidia.setGPUMemory(TOTAL_MIB)
            devs += [nvidia.Device(ID=device.fake_device_id(UUIDS[i], j)) for j in range(nvidia.getGPUMemory())]
        return devs, names

    monkeypatch.setattr(nvidia, "getDevices", getDevices)
    monkeypatch.setattr(device, "health_start", lambda *a, **k: None)
    # the event queue itself needs no driver: keep the real stop (it wakes blocked waiters)
    # drain anything a previous test left in the C-side event queue
    while device.health_wait(0) is not None:
        pass
