"""Interop of the native HTTP/2 + HPACK layer (csrc/daemon/h2.hpp) with grpcio's C-core stack:
unary, large responses through both flow-control windows, server streaming, many concurrent streams,
cancellation, error status, and the one-shot client against a grpcio server (what Register uses)."""
import os
import subprocess
import time
from concurrent import futures

import grpc
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.environ.get("H2_SELFTEST_BINARY") or os.path.join(ROOT, "build", "h2_selftest")

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


# ---- HPACK against the RFC's own vectors, and raw frames a C-core client never produces ---------------

def hpack(*blocks):
    out = subprocess.run([BIN, "hpack", *blocks], capture_output=True, text=True, timeout=10)
    return out.returncode, [b.strip().split("\n") if b.strip() else [] for b in out.stdout.split("--\n")[:-1]]


def test_hpack_rfc7541_appendix_c_request_examples():
    # C.3 (no Huffman) and C.4 (Huffman): three requests on one connection, dynamic table carried across
    rc, got = hpack("828684410f7777772e6578616d706c652e636f6d",
                    "828684be58086e6f2d6361636865",
                    "828785bf400a637573746f6d2d6b65790c637573746f6d2d76616c7565")
    assert rc == 0
    assert got[0] == [":method: GET", ":scheme: http", ":path: /", ":authority: www.example.com"]
    assert got[1] == [":method: GET", ":scheme: http", ":path: /", ":authority: www.example.com", "cache-control: no-cache"]
    assert got[2] == [":method: GET", ":scheme: https", ":path: /index.html", ":authority: www.example.com",
                      "custom-key: custom-value"]
    rc, hgot = hpack("828684418cf1e3c2e5f23a6ba0ab90f4ff",
                     "828684be5886a8eb10649cbf",
                     "828785bf408825a849e95ba97d7f8925a849e95bb8e8b4bf")
    assert rc == 0 and hgot == got
    # never-indexed literal, table-size update to 0 (evicts everything) then a stale index must fail
    rc, g = hpack("100870617373776f726406736563726574")
    assert rc == 0 and g[0] == ["password: secret"]
    rc, _ = hpack("828684410f7777772e6578616d706c652e636f6d", "20be")
    assert rc == 1


def frame(ftype, flags, stream, payload=b""):
    return len(payload).to_bytes(3, "big") + bytes([ftype, flags]) + stream.to_bytes(4, "big") + payload


def read_frames(sock, want_stream_end):
    import socket as _s
    sock.settimeout(5)
    buf, frames = b"", []
    while True:
        while len(buf) < 9:
            chunk = sock.recv(65536)
            if not chunk:
                return frames
            buf += chunk
        ln = int.from_bytes(buf[:3], "big")
        while len(buf) < 9 + ln:
            buf += sock.recv(65536)
        f = (buf[3], buf[4], int.from_bytes(buf[5:9], "big") & 0x7FFFFFFF, buf[9:9 + ln])
        buf = buf[9 + ln:]
        frames.append(f)
        if f[0] == 1 and f[2] == want_stream_end and f[1] & 0x1:  # HEADERS + END_STREAM = trailers
            return frames


def test_raw_frames_continuation_padding_priority(server):
    import socket as _s
    _, sock_path, _ = server
    s = _s.socket(_s.AF_UNIX, _s.SOCK_STREAM)
    s.connect(sock_path)
    s.sendall(b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n" + frame(4, 0, 0, (4).to_bytes(2, "big") + (1 << 20).to_bytes(4, "big")))
    # header block: indexed POST/http, literal :path (no huffman), content-type, te — split over HEADERS+2 CONTINUATIONs,
    # the HEADERS frame padded and carrying a PRIORITY section
    path = b"/test.Echo/Unary"
    block = bytes([0x83, 0x86, 0x04, len(path)]) + path + bytes([0x0f, 0x10, 16]) + b"application/grpc" + \
        bytes([0x00, 2]) + b"te" + bytes([8]) + b"trailers"
    a, b, c = block[:5], block[5:17], block[17:]
    padded = bytes([3]) + bytes([0x80, 0, 0, 0, 200]) + a + b"\0\0\0"   # pad length 3, exclusive dep on 0, weight 200
    s.sendall(frame(1, 0x8 | 0x20, 1, padded) + frame(9, 0, 1, b) + frame(9, 0x4, 1, c))
    msg = b"padded and split"
    grpc_msg = b"\0" + len(msg).to_bytes(4, "big") + msg
    s.sendall(frame(0, 0x8, 1, bytes([2]) + grpc_msg[:7] + b"\0\0") + frame(6, 0, 0, b"pingpong") +
              frame(0, 0x1, 1, grpc_msg[7:]))
    frames = read_frames(s, 1)
    assert any(f[0] == 6 and f[1] & 1 and f[3] == b"pingpong" for f in frames)  # PING ack, same payload
    assert any(f[0] == 4 and f[1] & 1 for f in frames)                         # SETTINGS ack
    data = b"".join(f[3] for f in frames if f[0] == 0 and f[2] == 1)
    assert data == grpc_msg
    trailers = [f for f in frames if f[0] == 1 and f[2] == 1 and f[1] & 1][0]
    assert b"grpc-status" in trailers[3] and trailers[3].endswith(b"\x010")
    # second request on the same connection, tiny peer window: the 100 KB answer must wait for WINDOW_UPDATEs
    s.sendall(frame(4, 0, 0, (4).to_bytes(2, "big") + (1000).to_bytes(4, "big")))
    block2 = bytes([0x83, 0x86, 0x04, 17]) + b"/test.Echo/Stream" + bytes([0x0f, 0x10, 16]) + b"application/grpc"
    req = b"1 100000"
    s.sendall(frame(1, 0x4, 3, block2) + frame(0, 0x1, 3, b"\0" + len(req).to_bytes(4, "big") + req))
    got, done = b"", False
    s.settimeout(5)
    buf = b""
    while not done:
        buf += s.recv(65536)
        while len(buf) >= 9:
            ln = int.from_bytes(buf[:3], "big")
            if len(buf) < 9 + ln:
                break
            t, fl, sid, pl = buf[3], buf[4], int.from_bytes(buf[5:9], "big"), buf[9:9 + ln]
            buf = buf[9 + ln:]
            if t == 0 and sid == 3:
                assert len(pl) <= 1000  # never more than the stream window we granted
                got += pl
                s.sendall(frame(8, 0, 3, len(pl).to_bytes(4, "big")) + frame(8, 0, 0, len(pl).to_bytes(4, "big")))
            if t == 1 and sid == 3 and fl & 1:
                done = True
    assert len(got) == 5 + 100000 and got[5:] == b"a" * 100000
    s.close()


# ---- JSON reader of the native daemon vs Python's json ---------------------------------------------------

import json as _json  # noqa: E402

from hypothesis import given, settings, strategies as st  # noqa: E402

_json_values = st.recursive(
    st.none() | st.booleans() | st.integers(-2**53, 2**53) | st.text(max_size=20) |
    st.floats(allow_nan=False, allow_infinity=False, width=32),
    lambda kids: st.lists(kids, max_size=4) | st.dictionaries(st.text(max_size=8), kids, max_size=4), max_leaves=20)


def native_json(text: bytes):
    out = subprocess.run([BIN, "json"], input=text, capture_output=True, timeout=10)
    return out.returncode, out.stdout


@settings(max_examples=150, deadline=None)
@given(doc=_json_values, ascii_only=st.booleans(), pretty=st.booleans())
def test_json_reader_agrees_with_python(doc, ascii_only, pretty):
    try:
        text = _json.dumps(doc, ensure_ascii=ascii_only, indent=2 if pretty else None).encode("utf-8")
    except UnicodeEncodeError:  # lone surrogates cannot be written as UTF-8
        return
    rc, out = native_json(text)
    assert rc == 0
    assert _json.loads(out.decode("utf-8", errors="surrogatepass")) == _json.loads(text)


@pytest.mark.parametrize("bad", [b"", b"{", b'{"a":}', b"[1,]", b'{"a" 1}', b'"unterminated', b"tru", b'{"a":1}x',
                                 b'"\\u12"', b'"\\q"', b"[" * 500 + b"]" * 500])
def test_json_reader_rejects_malformed(bad):
    rc, out = native_json(bad)
    assert rc == 1 and out.strip() == b"INVALID"


def test_server_survives_garbage_and_truncated_frames(server):
    """Protocol violations close that connection only; the server keeps serving others."""
    import random
    import socket as _s
    ch, sock_path, proc = server
    rng = random.Random(7)
    preface = b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n"
    blobs = [b"GET / HTTP/1.1\r\n\r\n", preface[:10], preface + b"\xff" * 64, preface + frame(1, 0x4, 1, b"\x80"),
             preface + frame(1, 0x4, 2, b"\x82"), preface + frame(0, 0, 1, b"data before headers"),
             preface + frame(9, 0x4, 1, b"orphan continuation"), preface + frame(4, 0, 0, b"\x00\x04\x00"),
             preface + frame(1, 0x8 | 0x4, 1, b"\xff\x82"), preface + frame(1, 0x4, 1, b"\x3f\xff\xff\xff\xff\xff\xff\xff\xff\xff\x7f"),
             preface + (200000).to_bytes(3, "big") + b"\x00\x00\x00\x00\x00\x01" + b"x" * 1000]
    blobs += [preface + bytes(rng.randrange(256) for _ in range(rng.randrange(1, 400))) for _ in range(40)]
    for b in blobs:
        s = _s.socket(_s.AF_UNIX, _s.SOCK_STREAM)
        s.settimeout(2)
        s.connect(sock_path)
        try:
            s.sendall(b)
            s.shutdown(_s.SHUT_WR)
            while s.recv(65536):
                pass
        except OSError:
            pass
        s.close()
        assert proc.poll() is None
    assert ch.unary_unary("/test.Echo/Unary")(b"still alive", timeout=5) == b"still alive"


def test_interop_with_curl_nghttp2(server, tmp_path):
    """A third HTTP/2 stack: curl + nghttp2 (h2c prior knowledge over the unix socket). nghttp2's HPACK encoder
    Huffman-codes and indexes differently from grpc's C-core."""
    import shutil
    if not shutil.which("curl") or "nghttp2" not in subprocess.run(["curl", "--version"], capture_output=True, text=True).stdout:
        pytest.skip("curl without nghttp2")
    _, sock_path, _ = server
    msg = b"hello from nghttp2 " * 50
    req = tmp_path / "req.bin"
    req.write_bytes(b"\0" + len(msg).to_bytes(4, "big") + msg)
    for _ in range(2):
        out = subprocess.run(["curl", "-sS", "--unix-socket", sock_path, "--http2-prior-knowledge", "-X", "POST",
                              "-H", "content-type: application/grpc", "-H", "te: trailers", "--data-binary", f"@{req}",
                              "-D", str(tmp_path / "hdr.txt"), "-o", str(tmp_path / "body.bin"),
                              "http://localhost/test.Echo/Unary"], capture_output=True, text=True, timeout=20)
        assert out.returncode == 0, out.stderr
        body = (tmp_path / "body.bin").read_bytes()
        assert body == b"\0" + len(msg).to_bytes(4, "big") + msg
        hdr = (tmp_path / "hdr.txt").read_text().lower()
        assert "http/2 200" in hdr and "content-type: application/grpc" in hdr and "grpc-status: 0" in hdr


# ---- limits: what grpc-go's defaults would also refuse ----------------------------------------------------

def _connect(sock_path):
    import socket as _s
    s = _s.socket(_s.AF_UNIX, _s.SOCK_STREAM)
    s.connect(sock_path)
    s.settimeout(10)
    s.sendall(b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n" + frame(4, 0, 0))
    return s


def _headers(path: bytes) -> bytes:
    return bytes([0x83, 0x86, 0x04, len(path)]) + path + bytes([0x0f, 0x10, 16]) + b"application/grpc"


def _frames_until(s, pred, limit=20000):
    buf, out = b"", []
    while len(out) < limit:
        try:
            chunk = s.recv(1 << 16)
        except OSError:
            break
        if not chunk:
            break
        buf += chunk
        while len(buf) >= 9:
            ln = int.from_bytes(buf[:3], "big")
            if len(buf) < 9 + ln:
                break
            f = (buf[3], buf[4], int.from_bytes(buf[5:9], "big") & 0x7FFFFFFF, buf[9:9 + ln])
            buf = buf[9 + ln:]
            out.append(f)
            if pred(f):
                return out
    return out


def test_message_over_4_mib_is_refused_not_buffered(server):
    """grpc.NewServer()'s default receive limit (vendor/google.golang.org/grpc/server.go:53): RESOURCE_EXHAUSTED."""
    _, sock_path, _ = server
    s = _connect(sock_path)
    big = (4 << 20) + 100
    s.sendall(frame(1, 0x4, 1, _headers(b"/test.Echo/Unary")))
    body = b"\0" + big.to_bytes(4, "big") + b"x" * big
    for off in range(0, len(body), 16384):
        chunk = body[off:off + 16384]
        s.sendall(frame(0, 0x1 if off + 16384 >= len(body) else 0, 1, chunk))
    frames = _frames_until(s, lambda f: f[0] == 1 and f[2] == 1 and f[1] & 1)
    trailers = [f for f in frames if f[0] == 1 and f[2] == 1 and f[1] & 1][0]
    assert trailers[3].endswith(b"grpc-status\x018")
    # the connection is still good for the next call
    msg = b"\0" + (2).to_bytes(4, "big") + b"ok"
    s.sendall(frame(1, 0x4, 3, _headers(b"/test.Echo/Unary")) + frame(0, 0x1, 3, msg))
    frames = _frames_until(s, lambda f: f[0] == 1 and f[2] == 3 and f[1] & 1)
    assert b"".join(f[3] for f in frames if f[0] == 0 and f[2] == 3) == msg
    s.close()


def test_continuation_flood_and_stream_id_reuse_close_the_connection(server):
    _, sock_path, _ = server
    s = _connect(sock_path)
    s.sendall(frame(1, 0, 1, _headers(b"/test.Echo/Unary")))  # no END_HEADERS: CONTINUATIONs follow, without end
    try:
        for _ in range(200):
            s.sendall(frame(9, 0, 1, bytes([0x00, 1, 0x61, 0x7f, 0x80, 0x7f]) + b"v" * 16000))
    except OSError:
        pass
    frames = _frames_until(s, lambda f: False)
    assert not any(f[0] == 1 for f in frames)  # never answered; the peer was dropped after 1 MiB of header block
    s.close()
    s = _connect(sock_path)
    msg = b"\0" + (2).to_bytes(4, "big") + b"ok"
    s.sendall(frame(1, 0x4, 5, _headers(b"/test.Echo/Unary")) + frame(0, 0x1, 5, msg))
    _frames_until(s, lambda f: f[0] == 1 and f[2] == 5 and f[1] & 1)
    s.sendall(frame(1, 0x4, 3, _headers(b"/test.Echo/Unary")) + frame(0, 0x1, 3, msg))  # ids must only grow
    frames = _frames_until(s, lambda f: False)
    assert not any(f[2] == 3 and f[0] in (0, 1) for f in frames)
    s.close()


def test_streams_beyond_the_advertised_maximum_are_refused(server):
    """SETTINGS_MAX_CONCURRENT_STREAMS = 1024: stream 1025 gets RST_STREAM(REFUSED_STREAM), the others finish."""
    _, sock_path, _ = server
    s = _connect(sock_path)
    n = 1024 + 6
    for i in range(n):  # open, do not finish: HEADERS only
        s.sendall(frame(1, 0x4, 2 * i + 1, _headers(b"/test.Echo/Unary")))
    msg = b"\0" + (2).to_bytes(4, "big") + b"ok"
    s.sendall(frame(0, 0x1, 1, msg))  # finish the first one
    frames = _frames_until(s, lambda f: f[0] == 1 and f[2] == 1 and f[1] & 1)
    refused = sorted(f[2] for f in frames if f[0] == 3 and f[3] == (7).to_bytes(4, "big"))
    assert refused == [2 * i + 1 for i in range(1024, n)]
    assert b"".join(f[3] for f in frames if f[0] == 0 and f[2] == 1) == msg
    s.close()
