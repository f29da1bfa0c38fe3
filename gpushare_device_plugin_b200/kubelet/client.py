"""Kubelet /pods/ client — pkg/kubelet/client/client.go kept as is in behaviour (north_star: "pkg/kubelet
client code is kept"): GET https://<addr>:<port>/pods/ with a bearer token or client cert, TLS
verification off (client.go:75-99), JSON -> v1.PodList (client.go:119-134). The response body is
closed here (the reference leaks it)."""
from __future__ import annotations

import http.client
import json
import ssl
from dataclasses import dataclass
from typing import Optional


@dataclass
class KubeletClientConfig:
    Address: str = "0.0.0.0"
    Port: int = 10250
    BearerToken: str = ""
    CertFile: str = ""
    KeyFile: str = ""
    HTTPTimeout: float = 10.0
    Scheme: str = "https"  # tests may use plain http against a loopback mock


class KubeletClient:
    def __init__(self, config: KubeletClientConfig):
        self.config = config
        self.ctx: Optional[ssl.SSLContext] = None
        if config.Scheme == "https":
            self.ctx = ssl.create_default_context()
            self.ctx.check_hostname = False  # Insecure: true (cmd/nvidia/main.go:40-41)
            self.ctx.verify_mode = ssl.CERT_NONE
            if config.CertFile and config.KeyFile:
                self.ctx.load_cert_chain(config.CertFile, config.KeyFile)

    def GetNodeRunningPods(self) -> dict:
        c = self.config
        if c.Scheme == "https":
            conn = http.client.HTTPSConnection(c.Address, c.Port, timeout=c.HTTPTimeout, context=self.ctx)
        else:
            conn = http.client.HTTPConnection(c.Address, c.Port, timeout=c.HTTPTimeout)
        try:
            headers = {"Authorization": "Bearer " + c.BearerToken} if c.BearerToken else {}
            conn.request("GET", "/pods/", headers=headers)
            r = conn.getresponse()
            body = r.read()
            if r.status != 200:
                raise RuntimeError(f"kubelet /pods/ returned {r.status}")
            return json.loads(body)
        finally:
            conn.close()


def NewKubeletClient(config: KubeletClientConfig) -> KubeletClient:
    return KubeletClient(config)
