"""Interop of the native HTTP/2 + HPACK layer (csrc/daemon/h2.hpp) with grpcio's C-core stack:
unary, large responses through both flow-control windows, server streaming, many concurrent streams,
cancellation, error status, and the one-shot client against a grpcio server (what Register uses)."""
import os
import subprocess
import threading
import time
from concurrent import futures

import grpc
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "h2_selftest")

pytestmark = pytest.mark.skipif(not os.access(BIN, os.X_OK), reason="build/h2_selftest not built")


@pytest.fixture
def server(tmp_path):
    sock = str(tmp_path / "h2.sock")
    p = subprocess.Popen([BIN, "serve", sock], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.stdout.readline().strip() == "ready"
    ch = grpc.insecure_channel("unix://" + sock, options=[("grpc.max_receive_message_length", 64 << 20)])
    grpc.channel_ready_future(ch).result(timeout=5)
    yield ch, sock, p
    ch.close()
    p.stdin.close()
    p.wait(timeout=5)


def test_unary_echo_and_status(server):
    ch, _, _ = server
    echo = ch.unary_unary("/test.Echo/Unary")
    for payload in (b"", b"x", b"hello" * 100, os.urandom(70000), os.urandom(1 << 20)):
        assert echo(payload, timeout=10) == payload
    with pytest.raises(grpc.RpcError) as e:
        ch.unary_unary("/test.Echo/Fail")(b"", timeout=5)
    assert e.value.code() == grpc.StatusCode.NOT_FOUND  # grpc-status 5
    with pytest.raises(grpc.RpcError) as e:
        ch.unary_unary("/test.Echo/Nope")(b"", timeout=5)
    assert e.value.code() == grpc.StatusCode.UNIMPLEMENTED


def test_streaming_through_flow_control(server):
    ch, _, _ = server
    msgs = list(ch.unary_stream("/test.Echo/Stream")(b"40 100000", timeout=30))  # 4 MB >> 64 KiB initial windows
    assert len(msgs) == 40 and all(len(m) == 100000 for m in msgs)
    assert msgs[3] == b"d" * 100000
    big = list(ch.unary_stream("/test.Echo/Stream")(b"2 3000000", timeout=30))
    assert [len(m) for m in big] == [3000000, 3000000]


def test_concurrent_streams_on_one_connection(server):
    ch, _, _ = server
    echo = ch.unary_unary("/test.Echo/Unary")

    def work(i):
        for j in range(50):
            p = bytes([i]) * (1 + (i * 37 + j) % 5000)
            assert echo(p, timeout=10) == p
    with futures.ThreadPoolExecutor(16) as ex:
        list(ex.map(work, range(16)))


def test_cancellation_reaches_the_handler(server):
    ch, _, p = server
    call = ch.unary_stream("/test.Echo/Forever")(b"")
    it = iter(call)
    assert next(it) == b"tick" and next(it) == b"tick"
    call.cancel()
    deadline = time.time() + 5
    line = ""
    while time.time() < deadline and "cancelled" not in line:
        line = p.stderr.readline()
    assert "forever: cancelled" in line
    # the connection is still usable
    assert ch.unary_unary("/test.Echo/Unary")(b"still here", timeout=5) == b"still here"


def test_native_client_against_grpcio_server(tmp_path):
    sock = str(tmp_path / "py.sock")
    seen = []

    def register(request, context):
        seen.append(request)
        return b"\x0a\x02ok"
    srv = grpc.server(futures.ThreadPoolExecutor(2))
    srv.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(
        "v1beta1.Registration", {"Register": grpc.unary_unary_rpc_method_handler(register)}),))
    srv.add_insecure_port("unix://" + sock)
    srv.start()
    try:
        req = bytes.fromhex("0a07763162657461311213616c6979756e67707573686172652e736f636b1a12616c6979756e2e636f6d2f6770752d6d656d")
        out = subprocess.run([BIN, "call", sock, "/v1beta1.Registration/Register", req.hex()], capture_output=True, text=True, timeout=10)
        assert out.returncode == 0, out.stdout + out.stderr
        status, resp = out.stdout.split()[:2]
        assert status == "0" and bytes.fromhex(resp) == b"\x0a\x02ok" and seen == [req]
        # unknown method -> UNIMPLEMENTED (12) comes back as a status, not a transport error
        out = subprocess.run([BIN, "call", sock, "/v1beta1.Registration/Nope", "00"], capture_output=True, text=True, timeout=10)
        assert out.stdout.split()[0] == "12"
    finally:
        srv.stop(0)
    out = subprocess.run([BIN, "call", sock, "/x/y", "00"], capture_output=True, text=True, timeout=10)
    assert out.stdout.split()[0] == "-1"  # nobody listening: dial error
