"""Stateful loopback mock of the kube-apiserver calls the path makes, plus the kubelet's /pods/.

  GET   /api/v1/nodes/<name>
  PATCH /api/v1/nodes/<name>/status            (strategic merge of status.capacity / allocatable)
  GET   /api/v1/pods?fieldSelector=spec.nodeName=<n>,status.phase=<p>
  PATCH /api/v1/namespaces/<ns>/pods/<name>    (strategic merge of metadata.annotations)
  GET   /pods/                                 (kubelet: every pod bound to the node)

The PATCH is applied, so an assigned pod stops matching the next Allocate (config 4's requirement).
`fail_next_patch(msg, times)` injects apiserver errors (OptimisticLock retry path).
"""
from __future__ import annotations

import copy
import json
import socket
import threading
import urllib.parse
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Dict, List, Optional


def make_node(name: str, gpu_count: Optional[int] = None, labels: Optional[dict] = None) -> dict:
    cap = {"cpu": "128", "memory": "2113929216Ki"}
    if gpu_count is not None:
        cap["aliyun.com/gpu-count"] = str(gpu_count)
    return {"kind": "Node", "apiVersion": "v1", "metadata": {"name": name, "labels": dict(labels or {})},
            "status": {"capacity": dict(cap), "allocatable": dict(cap)}}


def make_pod(i: int, node: str, gpu_mem: int = 4, idx: Optional[int] = 0, assume_time: Optional[int] = None,
             assigned: Optional[str] = "false", phase: str = "Pending", namespace: str = "default",
             containers: int = 1) -> dict:
    ann = {}
    if idx is not None:
        ann["ALIYUN_COM_GPU_MEM_IDX"] = str(idx)
    if assume_time is not None:
        ann["ALIYUN_COM_GPU_MEM_ASSUME_TIME"] = str(assume_time)
    if assigned is not None:
        ann["ALIYUN_COM_GPU_MEM_ASSIGNED"] = assigned
    per = gpu_mem // containers
    cs = [{"name": f"c{k}", "image": "busybox",
           "resources": {"limits": {"aliyun.com/gpu-mem": str(per if k else gpu_mem - per * (containers - 1))}}}
          for k in range(containers)]
    return {"kind": "Pod", "apiVersion": "v1",
            "metadata": {"name": f"pod-{i:02d}", "namespace": namespace, "uid": f"uid-{namespace}-{i:05d}",
                         "annotations": ann},
            "spec": {"nodeName": node, "containers": cs},
            "status": {"phase": phase}}


def config4_pods(node: str = "b200-0", n: int = 64, per_gpu: int = 8, mod: bool = False) -> List[dict]:
    """SURVEY.md §8(d) config 4 (IDX = i div 8) / config 5 (mod=True: IDX = i mod 8)."""
    return [make_pod(i, node, gpu_mem=4, idx=(i % 8 if mod else i // per_gpu),
                     assume_time=1_700_000_000_000_000_000 + i) for i in range(n)]


class MockKube:
    def __init__(self, node: dict, pods: List[dict], chunked_lists: bool = False, tls=None, client_ca=None):
        """tls = (certfile, keyfile) serves HTTPS; client_ca = CA file makes a client certificate mandatory."""
        self.chunked_lists = chunked_lists  # answer pod LISTs with Transfer-Encoding: chunked, like the apiserver
        self.auth_headers: List[Optional[str]] = []
        self.rv = 1000                      # cluster resourceVersion
        self.events: List[tuple] = []       # (rv, type, pod snapshot) — the watch cache
        self.watch_cv = threading.Condition()
        self.closing = False
        self.watches_served = 0
        self.enable_watch = True
        self.patch_delay = 0.0
        self.patched_ok: List[str] = []  # pod PATCHes that were applied (requests[] also holds the refused ones)
        self.required_token: Optional[str] = None  # when set: any other bearer token is answered 401 Unauthorized
        self.lock = threading.Lock()
        self.nodes: Dict[str, dict] = {node["metadata"]["name"]: node}
        self.pods: Dict[tuple, dict] = {(p["metadata"]["namespace"], p["metadata"]["name"]): p for p in pods}
        self.order = [(p["metadata"]["namespace"], p["metadata"]["name"]) for p in pods]
        self.requests: List[tuple] = []
        self._fail_patch: List[str] = []
        self.fail_lists = 0
        mock = self

        class H(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, *a):
                pass

            def setup(self):
                super().setup()
                self.request.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)

            def _send(self, code: int, obj, chunked: bool = False):
                body = obj if isinstance(obj, bytes) else \
                    json.dumps(obj, separators=(",", ":"), ensure_ascii=not chunked).encode()
                if chunked:
                    out = [b"HTTP/1.1 200 OK\r\nContent-Type: application/json\r\nTransfer-Encoding: chunked\r\n\r\n"]
                    for i in range(0, len(body), 1000):  # uneven chunks that split UTF-8 sequences and tokens
                        piece = body[i:i + 1000]
                        out.append(b"%x\r\n" % len(piece) + piece + b"\r\n")
                    out.append(b"0\r\n\r\n")
                    self.wfile.write(b"".join(out))
                    return
                head = (f"HTTP/1.1 {code} X\r\nContent-Type: application/json\r\nContent-Length: {len(body)}\r\n"
                        "\r\n").encode()
                self.wfile.write(head + body)  # one write: headers and body in the same segment

            def _status(self, code: int, message: str):
                self._send(code, {"kind": "Status", "apiVersion": "v1", "status": "Failure", "message": message,
                                  "code": code})

            def _authorized(self) -> bool:
                if mock.required_token is None or self.headers.get("Authorization") == "Bearer " + mock.required_token:
                    return True
                n = int(self.headers.get("Content-Length", "0"))
                if n:
                    self.rfile.read(n)
                self._status(401, "Unauthorized")
                return False

            def do_GET(self):
                mock.auth_headers.append(self.headers.get("Authorization"))
                if not self._authorized():
                    return
                u = urllib.parse.urlparse(self.path)
                parts = [p for p in u.path.split("/") if p]
                with mock.lock:
                    mock.requests.append(("GET", self.path))
                    if u.path == "/pods/" or u.path == "/pods":
                        return self._send(200, {"kind": "PodList", "apiVersion": "v1",
                                                "items": [copy.deepcopy(mock.pods[k]) for k in mock.order]})
                    if parts[:3] == ["api", "v1", "nodes"] and len(parts) == 3:
                        return self._send(200, {"kind": "NodeList", "apiVersion": "v1",
                                                "items": [copy.deepcopy(n) for n in mock.nodes.values()]})
                    if parts[:3] == ["api", "v1", "nodes"] and len(parts) == 4:
                        n = mock.nodes.get(parts[3])
                        return self._send(200, n) if n else self._status(404, f'nodes "{parts[3]}" not found')
                    if parts[:3] == ["api", "v1", "pods"] and urllib.parse.parse_qs(u.query).get("watch", ["0"])[0] in ("1", "true"):
                        if not mock.enable_watch:
                            return self._status(400, "watch is disabled on this mock")
                        q = urllib.parse.parse_qs(u.query)
                        sel = dict(kv.split("=", 1) for kv in q.get("fieldSelector", [""])[0].split(",") if "=" in kv)
                        since = int(q.get("resourceVersion", ["0"])[0] or 0)
                        mock.watches_served += 1
                        watch_args = (sel, since)
                    else:
                        watch_args = None
                    if watch_args is None and parts[:3] == ["api", "v1", "pods"]:
                        if mock.fail_lists > 0:
                            mock.fail_lists -= 1
                            return self._status(500, "etcdserver: request timed out")
                        sel = dict(kv.split("=", 1) for kv in
                                   urllib.parse.parse_qs(u.query).get("fieldSelector", [""])[0].split(",") if "=" in kv)
                        items = []
                        for k in mock.order:
                            p = mock.pods[k]
                            if "spec.nodeName" in sel and p["spec"].get("nodeName") != sel["spec.nodeName"]:
                                continue
                            if "status.phase" in sel and p["status"].get("phase") != sel["status.phase"]:
                                continue
                            items.append(copy.deepcopy(p))
                        return self._send(200, {"kind": "PodList", "apiVersion": "v1",
                                                "metadata": {"resourceVersion": str(mock.rv)}, "items": items},
                                          chunked=mock.chunked_lists)
                if watch_args is not None:
                    return self._watch(*watch_args)
                self._status(404, "not found")

            def _watch(self, sel, since):
                """GET /api/v1/pods?watch=true: chunked stream of {"type","object"} lines from resourceVersion `since`.
                A pod that stops matching the field selector is reported as DELETED, as the apiserver does."""
                def matches(node, phase):
                    return (("spec.nodeName" not in sel or node == sel["spec.nodeName"]) and
                            ("status.phase" not in sel or phase == sel["status.phase"]))
                try:
                    self.wfile.write(b"HTTP/1.1 200 OK\r\nContent-Type: application/json\r\nTransfer-Encoding: chunked\r\n\r\n")
                    self.wfile.flush()
                    cursor = 0
                    while True:
                        with mock.watch_cv:
                            while cursor >= len(mock.events) and not mock.closing:
                                mock.watch_cv.wait(0.2)
                            if mock.closing:
                                break
                            batch = mock.events[cursor:]
                            cursor = len(mock.events)
                        for rv, etype, body, node, phase in batch:
                            if rv <= since:
                                continue
                            if etype == "ERROR":  # e.g. 410 Gone: one Status line, then the server ends the stream
                                line = b'{"type":"ERROR","object":' + body + b"}\n"
                                self.wfile.write(b"%x\r\n" % len(line) + line + b"\r\n0\r\n\r\n")
                                self.wfile.flush()
                                self.close_connection = True
                                return
                            if not matches(node, phase):
                                etype = "DELETED"
                            line = b'{"type":"' + etype.encode() + b'","object":' + body + b"}\n"
                            self.wfile.write(b"%x\r\n" % len(line) + line + b"\r\n")
                            self.wfile.flush()
                    self.wfile.write(b"0\r\n\r\n")
                except OSError:
                    pass
                self.close_connection = True

            def do_PATCH(self):
                mock.auth_headers.append(self.headers.get("Authorization"))
                if not self._authorized():
                    return
                u = urllib.parse.urlparse(self.path)
                parts = [p for p in u.path.split("/") if p]
                body = self.rfile.read(int(self.headers.get("Content-Length", "0")))
                if mock.patch_delay:  # a slow apiserver: the write lands this much later than it was sent
                    threading.Event().wait(mock.patch_delay)  # not time.sleep: tests stub that out
                with mock.lock:
                    mock.requests.append(("PATCH", self.path, body, self.headers.get("Content-Type")))
                    try:
                        patch = json.loads(body)
                    except ValueError:
                        return self._status(400, "invalid JSON patch")
                    if parts[:3] == ["api", "v1", "nodes"] and len(parts) == 5 and parts[4] == "status":
                        n = mock.nodes.get(parts[3])
                        if not n:
                            return self._status(404, f'nodes "{parts[3]}" not found')
                        for k in ("capacity", "allocatable"):
                            n["status"].setdefault(k, {}).update((patch.get("status") or {}).get(k) or {})
                        return self._send(200, n)
                    if parts[:3] == ["api", "v1", "namespaces"] and len(parts) == 6 and parts[4] == "pods":
                        if mock._fail_patch:
                            return self._status(409, mock._fail_patch.pop(0))
                        p = mock.pods.get((parts[3], parts[5]))
                        if not p:
                            return self._status(404, f'pods "{parts[5]}" not found')
                        p["metadata"].setdefault("annotations", {}).update(
                            (patch.get("metadata") or {}).get("annotations") or {})
                        mock.patched_ok.append(self.path)
                        return self._send(200, mock._emit("MODIFIED", p))
                self._status(404, "not found")

        class Server(ThreadingHTTPServer):
            request_queue_size = 1024  # default 5: a burst of new client connections would hit SYN retransmits (1 s)

        self.httpd = Server(("127.0.0.1", 0), H)
        self.scheme = "http"
        if tls is not None:
            import ssl
            ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
            ctx.load_cert_chain(*tls)
            if client_ca is not None:
                ctx.verify_mode = ssl.CERT_REQUIRED
                ctx.load_verify_locations(client_ca)
            self.httpd.socket = ctx.wrap_socket(self.httpd.socket, server_side=True)
            self.scheme = "https"
        self.httpd.daemon_threads = True
        self.port = self.httpd.server_address[1]
        self.url = f"{self.scheme}://127.0.0.1:{self.port}"
        self.thread = threading.Thread(target=self.httpd.serve_forever, name="mock-kube", daemon=True)
        self.thread.start()

    def _emit(self, etype: str, pod: dict) -> bytes:
        """caller holds self.lock; the object is serialised once, for the event log and for the caller's response"""
        self.rv += 1
        pod.setdefault("metadata", {})["resourceVersion"] = str(self.rv)
        body = json.dumps(pod, separators=(",", ":")).encode()
        with self.watch_cv:
            self.events.append((self.rv, etype, body, (pod.get("spec") or {}).get("nodeName"),
                                (pod.get("status") or {}).get("phase")))
            self.watch_cv.notify_all()
        return body

    def add_pod(self, pod: dict):
        with self.lock:
            key = (pod["metadata"]["namespace"], pod["metadata"]["name"])
            self.pods[key] = pod
            self.order.append(key)
            self._emit("ADDED", pod)

    def delete_pod(self, name: str, namespace: str = "default"):
        with self.lock:
            pod = self.pods.pop((namespace, name))
            self.order.remove((namespace, name))
            self._emit("DELETED", pod)

    def expire_watches(self):
        """Every open watch gets an ERROR event (410 Gone, "too old resource version") and is closed by the server."""
        with self.lock:
            self.rv += 1
            status = {"kind": "Status", "apiVersion": "v1", "status": "Failure", "reason": "Expired", "code": 410,
                      "message": "too old resource version"}
            with self.watch_cv:
                self.events.append((self.rv, "ERROR", json.dumps(status).encode(), None, None))
                self.watch_cv.notify_all()

    def fail_next_patch(self, message: str, times: int = 1):
        with self.lock:
            self._fail_patch.extend([message] * times)

    def pod(self, name: str, namespace: str = "default") -> dict:
        with self.lock:
            return copy.deepcopy(self.pods[(namespace, name)])

    def close(self):
        with self.watch_cv:
            self.closing = True
            self.watch_cv.notify_all()
        self.httpd.shutdown()
        self.httpd.server_close()


def main(argv=None):
    """Stand-alone mock (its own process, so it does not share a GIL with the plugin under test):
    python -m gpushare_device_plugin_b200.testing.mock_kube --node b200-0 --pods 1024 [--mod]
    prints the port on stdout, serves until stdin closes."""
    import argparse
    import sys
    ap = argparse.ArgumentParser()
    ap.add_argument("--node", default="b200-0")
    ap.add_argument("--pods", type=int, default=64)
    ap.add_argument("--mod", action="store_true")
    a = ap.parse_args(argv)
    m = MockKube(make_node(a.node, gpu_count=8), config4_pods(a.node, a.pods, mod=a.mod))
    print(m.port, flush=True)
    sys.stdin.read()
    m.close()


if __name__ == "__main__":
    main()
