"""Fake kubelet: serves v1beta1.Registration/Register on <dir>/kubelet.sock (recording the raw
request bytes) and acts as the device manager's client of the plugin socket (raw-bytes RPCs)."""
from __future__ import annotations

import os
import queue
from concurrent import futures
from typing import Iterator

import grpc


class FakeKubelet:
    def __init__(self, plugin_dir: str):
        self.dir = plugin_dir
        self.socket = os.path.join(plugin_dir, "kubelet.sock")
        self.register_requests: "queue.Queue[bytes]" = queue.Queue()
        self.server = None
        self.start()

    def start(self):
        try:
            os.remove(self.socket)
        except FileNotFoundError:
            pass

        def register(request: bytes, context) -> bytes:
            self.register_requests.put(request)
            return b""  # Empty

        h = grpc.method_handlers_generic_handler(
            "v1beta1.Registration", {"Register": grpc.unary_unary_rpc_method_handler(register)})
        self.server = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
        self.server.add_generic_rpc_handlers((h,))
        self.server.add_insecure_port("unix://" + self.socket)
        self.server.start()

    def stop(self):
        if self.server:
            # wait for the teardown: grpc unlinks the unix socket when the listener is destroyed, which would otherwise
            # remove the file a kubelet restarted on the same path has just bound
            self.server.stop(0).wait(10)
            self.server = None
        try:
            os.remove(self.socket)
        except FileNotFoundError:
            pass

    # ---- device-manager side: talk to the plugin -------------------------------------------
    def channel(self, endpoint: str) -> grpc.Channel:
        ch = grpc.insecure_channel("unix://" + os.path.join(self.dir, endpoint))
        grpc.channel_ready_future(ch).result(timeout=5)
        return ch

    @staticmethod
    def get_options(ch) -> bytes:
        return ch.unary_unary("/v1beta1.DevicePlugin/GetDevicePluginOptions")(b"", timeout=5)

    @staticmethod
    def pre_start(ch) -> bytes:
        return ch.unary_unary("/v1beta1.DevicePlugin/PreStartContainer")(b"", timeout=5)

    @staticmethod
    def list_and_watch(ch) -> Iterator[bytes]:
        return ch.unary_stream("/v1beta1.DevicePlugin/ListAndWatch")(b"")

    @staticmethod
    def allocate(ch, request: bytes, timeout: float = 30) -> bytes:
        return ch.unary_unary("/v1beta1.DevicePlugin/Allocate")(request, timeout=timeout)
