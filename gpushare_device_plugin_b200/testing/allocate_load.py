"""Load generator for Allocate(): N client threads, each with its own
channel to a plugin socket, issuing 4-ID requests; returns latencies in microseconds."""
from __future__ import annotations

import threading
import time
from typing import List

import grpc


def allocate_request(ids: List[str]) -> bytes:
    """AllocateRequest{container_requests:[{devicesIDs: ids}]} (v1beta1/api.proto:112-122)."""
    inner = b"".join(b"\x0a" + bytes([len(s)]) + s.encode() for s in ids)
    assert len(inner) < 16384
    ln = len(inner)
    var = bytes([ln]) if ln < 128 else bytes([(ln & 0x7F) | 0x80, ln >> 7])
    return b"\x0a" + var + inner


def run(socket_path: str, concurrency: int, total: int, uuids: List[str]) -> dict:
    per = [total // concurrency + (1 if i < total % concurrency else 0) for i in range(concurrency)]
    lat: List[List[float]] = [[] for _ in range(concurrency)]
    errs = [0] * concurrency
    chans = [grpc.insecure_channel("unix://" + socket_path) for _ in range(concurrency)]
    for ch in chans:
        grpc.channel_ready_future(ch).result(timeout=10)
    calls = [ch.unary_unary("/v1beta1.DevicePlugin/Allocate") for ch in chans]
    start = threading.Barrier(concurrency + 1)

    def worker(i: int):
        req = allocate_request([f"{uuids[i % len(uuids)]}-_-{j}" for j in range(4)])
        start.wait()
        for _ in range(per[i]):
            t0 = time.perf_counter_ns()
            resp = calls[i](req, timeout=120)
            lat[i].append((time.perf_counter_ns() - t0) / 1e3)
            if b"no-gpu-has" in resp:
                errs[i] += 1

    ts = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(concurrency)]
    for t in ts:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in ts:
        t.join()
    wall = time.perf_counter() - t0
    for ch in chans:
        ch.close()
    flat = sorted(x for l in lat for x in l)
    pick = lambda q: flat[min(len(flat) - 1, int(q * len(flat)))]  # noqa: E731
    return {"concurrency": concurrency, "requests": len(flat), "p50_us": pick(0.5), "p99_us": pick(0.99),
            "mean_us": sum(flat) / len(flat), "req_per_s": len(flat) / wall, "error_responses": sum(errs)}


if __name__ == "__main__":  # own process: python -m ...allocate_load <socket> <concurrency> <total> <uuid,uuid,...>
    import json
    import sys
    print(json.dumps(run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4].split(","))))
