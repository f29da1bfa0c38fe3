"""GPU-box parity of the inventory half against the reference's own NVML path: the C restatement
linked with the reference's nvml_dl.c (oracle/_ref/ref_inventory) is run in the same process tree,
on the same devices, and every field the plugin consumes must agree bit for bit — uuid, minor,
total bytes, MiB, slice count, device order and the ListAndWatchResponse bytes."""
import json
import os
import subprocess

import pytest

from oracle import wire_oracle as wo

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_inventory")


def run_ref(tmp_path, unit="GiB"):
    lw = tmp_path / f"lw_{unit}.bin"
    out = subprocess.run([REF_BIN, "inventory", "--unit", unit, "--lw-out", str(lw)], check=True,
                         capture_output=True, text=True)
    return json.loads(out.stdout), lw.read_bytes()


def test_reference_binary_is_the_reference_shim():
    assert os.access(REF_BIN, os.X_OK), "oracle/_ref/ref_inventory must be prebuilt (make -C oracle) and shipped"


def test_inventory_matches_reference_nvml_path(gsb, tmp_path):
    ref, ref_lw = run_ref(tmp_path)
    n = gsb.device_count()
    assert n == ref["n_gpus"] >= 1
    infos = [gsb.device_info(i) for i in range(n)]
    for mine, theirs in zip(infos, ref["devices"]):
        assert mine.uuid == theirs["uuid"]
        assert mine.minor == theirs["minor"]
        assert f"/dev/nvidia{mine.minor}" == theirs["path"]
        assert mine.total_bytes == theirs["total_bytes"]
        assert mine.total_mib == theirs["memory_mib"]
        assert mine.bus_id.lower() == theirs["bus_id"].lower()
        assert mine.cc == (10, 0) and mine.sm_count == 148
    # the reference derives the process-global slice count from device 0 only (nvidia.go:70-72)
    s = gsb.slices(infos[0].total_mib, True)
    assert s == ref["gpu_memory"] == infos[0].total_bytes >> 30
    lw = gsb.encode_list_and_watch([i.uuid for i in infos], s)
    assert lw == ref_lw
    assert len(lw) == ref["lw_len"]
    # ... and both equal the pure-Python restatement of getDevices + gogo marshal
    devs, name_map, mem = wo.getDevices([{"uuid": d["uuid"], "path": d["path"], "memory_mib": d["memory_mib"]}
                                         for d in ref["devices"]])
    assert mem == s and wo.marshal_ListAndWatchResponse(devs) == lw
    assert name_map == {i.uuid: i.minor for i in infos}
    assert gsb.fake_device_id(infos[0].uuid, 0) == ref["first_id"].replace(ref["devices"][0]["uuid"], infos[0].uuid)
    assert gsb.fake_device_id(infos[-1].uuid, s - 1) == ref["last_id"]


def test_b200_advertises_179_slices(gsb):
    info = gsb.device_info(0)
    # nvidia-smi's 183359 MiB; cuDeviceTotalMem alone would give 178 (profiles/envprobe_r01.txt)
    assert info.total_mib == 183359 and gsb.slices(info.total_mib, True) == 179
    assert info.cuda_total_bytes >> 30 == 178
    assert len(gsb.encode_list_and_watch([info.uuid], 179)) == 10451


def test_mib_unit_matches_reference(gsb, tmp_path):
    ref, _ = run_ref(tmp_path, "MiB")
    info = gsb.device_info(0)
    assert gsb.slices(info.total_mib, False) == ref["gpu_memory"] == 183359
