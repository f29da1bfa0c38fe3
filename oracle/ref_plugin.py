"""Reference-behaviour device-plugin server — CPU BASELINE / TEST INFRASTRUCTURE ONLY.

A runnable restatement of pkg/gpu/nvidia/{server.go,allocate.go,podmanager.go} for the Allocate()
baseline of SURVEY.md §8(d) configs 4-5, keeping the reference's performance-relevant behaviours:
  * one global mutex held across ALL network I/O (allocate.go:59-60)
  * a full LIST of the node's pending pods on every call (podmanager.go:142-160) + JSON decode
  * gogo-style decode/encode of the request/response (oracle/wire_oracle.py)
  * >= 6 synchronous log lines per call at the DaemonSet's --v=5 (allocate.go:46,57,61,80,111,190-191)
  * strategic-merge PATCH of the matched pod, one retry on the optimistic-lock message (:135-149)
It shares no code with the product (own HTTP calls, own codec); it is served over the same grpcio
transport so that only the plugin logic differs between the two arms. Never imported by the product.
"""
from __future__ import annotations

import http.client
import json
import logging
import os
import socket
import threading
import urllib.parse
from concurrent import futures

import grpc

from . import wire_oracle as wo

log = logging.getLogger("ref_plugin")


class RefPlugin:
    def __init__(self, api_url: str, node: str, dev_name_map: dict, gpu_memory: int, socket: str,
                 metric: str = wo.GiBPrefix, max_workers: int = 64, log_path: str = os.devnull):
        u = urllib.parse.urlparse(api_url)
        self.host, self.port = u.hostname, u.port
        self.node, self.devNameMap, self.gpuMemory, self.metric = node, dev_name_map, gpu_memory, metric
        self.socket = socket
        self.mu = threading.Lock()
        self._tl = threading.local()
        h = logging.FileHandler(log_path)  # glog writes synchronously; so does this handler
        h.setFormatter(logging.Formatter("I%(asctime)s %(message)s"))
        log.handlers[:] = [h]
        log.setLevel(logging.INFO)
        log.propagate = False
        self.server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
        self.server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler("v1beta1.DevicePlugin", {
            "Allocate": grpc.unary_unary_rpc_method_handler(self.Allocate)}),))
        self.server.add_insecure_port("unix://" + socket)

    def start(self):
        self.server.start()

    def stop(self):
        self.server.stop(0)
        try:
            os.remove(self.socket)
        except FileNotFoundError:
            pass

    def _http(self, method, path, body=None, ctype=None):
        c = getattr(self._tl, "c", None)
        if c is None:
            c = self._tl.c = http.client.HTTPConnection(self.host, self.port, timeout=30)
            c.connect()
            c.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)  # as Go's net/http does
        try:
            c.request(method, path, body=body, headers={"Content-Type": ctype} if ctype else {})
            r = c.getresponse()
            data = r.read()
        except Exception:
            c.close()
            self._tl.c = None
            raise
        if r.status >= 400:
            raise RuntimeError(json.loads(data).get("message", "error"))
        return json.loads(data)

    def Allocate(self, request: bytes, context) -> bytes:
        reqs = [[v2.decode() for f2, w2, v2 in wo._fields(v) if f2 == 1] for f, w, v in wo._fields(request) if f == 1]
        log.info("----Allocating GPU for gpu mem is started----")
        log.info("RequestPodGPUs: %d", sum(len(r) for r in reqs))
        with self.mu:
            log.info("checking...")
            sel = urllib.parse.urlencode({"fieldSelector": f"spec.nodeName={self.node},status.phase=Pending"})
            try:
                pods = self._http("GET", f"/api/v1/pods?{sel}")["items"]
            except Exception:
                return wo.marshal_AllocateResponse(wo.buildErrResponse(reqs, sum(map(len, reqs)), self.metric, self.gpuMemory))
            for p in pods:  # podmanager.go:184-201: one log line per listed pod
                log.info("list pod %s in ns %s in node %s and status is %s", p["metadata"]["name"],
                         p["metadata"]["namespace"], self.node, p["status"]["phase"])

            def patch(pod, body):
                try:
                    self._http("PATCH", f"/api/v1/namespaces/{pod['metadata']['namespace']}/pods/{pod['metadata']['name']}",
                               body, "application/strategic-merge-patch+json")
                    return None
                except Exception as e:  # noqa: BLE001
                    return str(e)
            envs, pod = wo.Allocate(reqs, pods, self.node, self.devNameMap, self.gpuMemory, self.metric, False, False, patch)
            if pod is not None:
                log.info("Found Assumed GPU shared Pod %s in ns %s with GPU Memory %d", pod["metadata"]["name"],
                         pod["metadata"]["namespace"], sum(map(len, reqs)))
                log.info("gpu index %s", envs[0].get(wo.EnvResourceIndex))
            log.info("pod %s, new allocated GPUs info %s", pod["metadata"]["name"] if pod else "", envs)
            log.info("----Allocating GPU for gpu mem for %s is ended----", pod["metadata"]["name"] if pod else "")
            return wo.marshal_AllocateResponse(envs)
