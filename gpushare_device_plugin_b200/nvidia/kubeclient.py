"""The four kube-apiserver calls the path makes (client-go in the reference): node GET, node-status
strategic-merge PATCH, pod LIST with a field selector, pod strategic-merge PATCH. Plain HTTPS/JSON
over one keep-alive connection per thread. Control-plane I/O: restated, not accelerated."""
from __future__ import annotations

import http.client
import json
import os
import socket
import ssl
import threading
import urllib.parse
from typing import Optional, Tuple

import yaml

SA_DIR = "/var/run/secrets/kubernetes.io/serviceaccount"


class ApiError(Exception):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status, self.message = status, message

    def __str__(self):  # client-go surfaces Status.message as err.Error()
        return self.message


class Clientset:
    def __init__(self, server: str, token: str = "", ca_file: Optional[str] = None, insecure: bool = False,
                 cert: Optional[Tuple[str, str]] = None, timeout: float = 30.0, ca_data: Optional[str] = None):
        u = urllib.parse.urlparse(server)
        self.scheme, self.host, self.port = u.scheme, u.hostname, u.port or (443 if u.scheme == "https" else 80)
        self.token, self.timeout = token, timeout
        self.ctx = None
        if self.scheme == "https":
            if ca_data and not insecure:
                self.ctx = ssl.create_default_context(cadata=ca_data)
            elif ca_file and not insecure:
                self.ctx = ssl.create_default_context(cafile=ca_file)
            else:
                self.ctx = ssl.create_default_context()
            if insecure:
                self.ctx.check_hostname = False
                self.ctx.verify_mode = ssl.CERT_NONE
            if cert:
                self.ctx.load_cert_chain(*cert)
        self._tl = threading.local()
        self.token_file: Optional[str] = None  # in-cluster only: the kubelet rewrites it before the token expires

    def _conn(self) -> http.client.HTTPConnection:
        c = getattr(self._tl, "conn", None)
        if c is None:
            if self.scheme == "https":
                c = http.client.HTTPSConnection(self.host, self.port, timeout=self.timeout, context=self.ctx)
            else:
                c = http.client.HTTPConnection(self.host, self.port, timeout=self.timeout)
            c.connect()
            # Go's net/http (client-go) disables Nagle; without this a PATCH (headers + body = two
            # small writes) stalls ~40 ms on the peer's delayed ACK
            c.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            self._tl.conn = c
        return c

    def request(self, method: str, path: str, body: Optional[bytes] = None, content_type: Optional[str] = None):
        try:
            return self._request(method, path, body, content_type)
        except ApiError as e:
            if e.status != 401 or not self.token_file:
                raise
            with open(self.token_file) as f:  # rotated service-account token: pick up the new one, once
                fresh = f.read().strip()
            if not fresh or fresh == self.token:
                raise
            self.token = fresh
            return self._request(method, path, body, content_type)

    def _request(self, method: str, path: str, body: Optional[bytes] = None, content_type: Optional[str] = None):
        headers = {"Accept": "application/json"}
        if self.token:
            headers["Authorization"] = "Bearer " + self.token
        if content_type:
            headers["Content-Type"] = content_type
        for attempt in (0, 1):
            c = self._conn()
            try:
                c.request(method, path, body=body, headers=headers)
                r = c.getresponse()
                data = r.read()
                break
            except (http.client.HTTPException, ConnectionError, OSError):
                c.close()
                self._tl.conn = None
                if attempt:
                    raise
        if r.status >= 400:
            try:
                msg = json.loads(data).get("message") or data.decode(errors="replace")
            except Exception:
                msg = data.decode(errors="replace")
            raise ApiError(r.status, msg)
        return json.loads(data) if data else {}

    # --- the calls podmanager.go / allocate.go make ---------------------------------------
    def get_node(self, name: str) -> dict:
        return self.request("GET", f"/api/v1/nodes/{name}")

    def patch_node_status(self, name: str, patch: dict) -> dict:
        return self.request("PATCH", f"/api/v1/nodes/{name}/status", json.dumps(patch, separators=(",", ":")).encode(),
                            "application/strategic-merge-patch+json")

    def list_pods(self, field_selector: str) -> dict:
        q = urllib.parse.urlencode({"fieldSelector": field_selector})
        return self.request("GET", f"/api/v1/pods?{q}")

    def patch_pod(self, namespace: str, name: str, body: bytes) -> dict:
        return self.request("PATCH", f"/api/v1/namespaces/{namespace}/pods/{name}", body,
                            "application/strategic-merge-patch+json")


def from_environment() -> Clientset:
    """kubeInit's config choice (podmanager.go:29-50): $KUBECONFIG if the file exists, else in-cluster."""
    path = os.environ.get("KUBECONFIG", "")
    if path and os.path.exists(path):
        with open(path) as f:
            cfg = yaml.safe_load(f)
        ctx_name = cfg.get("current-context")
        ctx = next((c["context"] for c in cfg.get("contexts", []) if c["name"] == ctx_name), None) or \
            cfg["contexts"][0]["context"]
        cluster = next(c["cluster"] for c in cfg["clusters"] if c["name"] == ctx["cluster"])
        user = next((u["user"] for u in cfg.get("users", []) if u["name"] == ctx.get("user")), {}) or {}
        import base64
        import tempfile
        cert = (user["client-certificate"], user["client-key"]) if "client-certificate" in user else None
        if "client-certificate-data" in user:  # ssl wants files: materialise the embedded PEMs privately
            files = []
            for key in ("client-certificate-data", "client-key-data"):
                f = tempfile.NamedTemporaryFile("wb", suffix=".pem", delete=False)
                os.chmod(f.name, 0o600)
                f.write(base64.b64decode(user[key]))
                f.close()
                files.append(f.name)
            cert = (files[0], files[1])
        ca_data = cluster.get("certificate-authority-data")
        return Clientset(cluster["server"], token=user.get("token", ""), ca_file=cluster.get("certificate-authority"),
                         insecure=bool(cluster.get("insecure-skip-tls-verify")), cert=cert,
                         ca_data=base64.b64decode(ca_data).decode() if ca_data else None)
    host, port = os.environ.get("KUBERNETES_SERVICE_HOST"), os.environ.get("KUBERNETES_SERVICE_PORT")
    if not host or not port:
        raise RuntimeError("unable to load in-cluster configuration, KUBERNETES_SERVICE_HOST and "
                           "KUBERNETES_SERVICE_PORT must be defined")
    with open(os.path.join(SA_DIR, "token")) as f:
        token = f.read().strip()
    cs = Clientset(f"https://{host}:{port}", token=token, ca_file=os.path.join(SA_DIR, "ca.crt"))
    cs.token_file = os.path.join(SA_DIR, "token")
    return cs
