"""python tools/fuzz_apiserver.py build/san/gsbd.asan <seconds> <seed>
gsbd (ASan) against an apiserver that answers with damaged HTTP: truncated bodies, bad chunk sizes, lying
Content-Length, garbage status lines, early closes. The daemon may refuse to start or refuse requests; it must not crash."""
import json, os, random, socket, subprocess, sys, tempfile, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpushare_device_plugin_b200.testing.fake_kubelet import FakeKubelet
from gpushare_device_plugin_b200.testing.mock_kube import config4_pods, make_node
from oracle import wire_oracle as wo
BIN, secs, seed = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
rnd = random.Random(seed)
NODE = "b200-0"
pods = config4_pods(NODE, 32)
def good(path, method):
    if "/nodes/" in path: body = make_node(NODE, gpu_count=8)
    elif method == "PATCH": body = pods[0]
    else: body = {"kind": "PodList", "apiVersion": "v1", "metadata": {"resourceVersion": "1000"}, "items": pods}
    return json.dumps(body).encode()
def damaged(body):
    r = rnd.random() if rnd.random() < 0.35 else 0.0
    head = b"HTTP/1.1 200 OK\r\nContent-Type: application/json\r\n"
    if r < 0.25: return head + b"Content-Length: %d\r\n\r\n" % len(body) + body           # fine
    if r < 0.35: return head + b"Content-Length: %d\r\n\r\n" % (len(body) + rnd.randint(1, 50)) + body  # lies, then closes
    if r < 0.45: return head + b"Content-Length: %d\r\n\r\n" % len(body) + body[: rnd.randint(0, len(body))]
    if r < 0.60:
        out = head + b"Transfer-Encoding: chunked\r\n\r\n"; i = 0
        while i < len(body):
            n = rnd.randint(1, 700); piece = body[i:i + n]; i += n
            size = b"%x" % len(piece)
            if rnd.random() < 0.05: size = rnd.choice([b"zz", b"-1", b"ffffffffffffffff", b"", b"7fffffff"])
            out += size + rnd.choice([b"\r\n", b";ext=1\r\n", b"\n"]) + piece + rnd.choice([b"\r\n", b"\r\n", b""])
        return out + rnd.choice([b"0\r\n\r\n", b"0\r\n", b"", b"0\r\nTrailer: x\r\n\r\n"])
    if r < 0.70: return rnd.choice([b"HTTP/1.1 9999999999 X\r\n\r\n", b"HTTP/1.1\r\n\r\n", b"\r\n\r\n", b"garbage", b"HTTP/1.1 200 OK\r\nContent-Length: -5\r\n\r\n", b"HTTP/1.1 200 OK\r\nContent-Length: 99999999999999999999\r\n\r\n"])
    if r < 0.80: return head + b"Content-Length: %d\r\n\r\n" % len(body) + bytes(rnd.getrandbits(8) for _ in range(len(body)))
    if r < 0.90:
        b2 = bytearray(body)
        for _ in range(rnd.randint(1, 8)): b2[rnd.randrange(len(b2))] = rnd.getrandbits(8)
        return head + b"Content-Length: %d\r\n\r\n" % len(b2) + bytes(b2)
    return b""
def serve(conn):
    conn.settimeout(2)
    try:
        buf = b""
        while True:
            while b"\r\n\r\n" not in buf:
                d = conn.recv(65536)
                if not d: return
                buf += d
            head, _, rest = buf.partition(b"\r\n\r\n")
            line = head.split(b"\r\n")[0].decode(errors="replace").split(" ")
            clen = 0
            for h in head.split(b"\r\n")[1:]:
                if h.lower().startswith(b"content-length:"): clen = int(h.split(b":")[1])
            while len(rest) < clen: rest += conn.recv(65536)
            buf = rest[clen:]
            if "watch=true" in line[1] and rnd.random() < 0.5:
                conn.sendall(b"HTTP/1.1 200 OK\r\nTransfer-Encoding: chunked\r\n\r\n")
                for _ in range(rnd.randint(0, 5)):
                    ev = json.dumps({"type": rnd.choice(["ADDED", "MODIFIED", "DELETED", "ERROR", "BOOKMARK", "x"]), "object": rnd.choice(pods)}).encode()
                    if rnd.random() < 0.3: ev = ev[: rnd.randint(0, len(ev))]
                    ev += b"\n"
                    conn.sendall(b"%x\r\n" % len(ev) + ev + b"\r\n"); time.sleep(0.01)
                return
            out = damaged(good(line[1], line[0]))
            conn.sendall(out)
            if rnd.random() < 0.5: return
    except (OSError, ValueError, IndexError):
        pass
    finally:
        conn.close()
ls = socket.socket(); ls.bind(("127.0.0.1", 0)); ls.listen(256); port = ls.getsockname()[1]
def acceptor():
    while True:
        try: c, _ = ls.accept()
        except OSError: return
        threading.Thread(target=serve, args=(c,), daemon=True).start()
threading.Thread(target=acceptor, daemon=True).start()
end = time.time() + secs; runs = registered = answered = 0; codes = {}
while time.time() < end:
    tmp = tempfile.mkdtemp(prefix="gsb-fz-")
    kubelet = FakeKubelet(tmp)
    env = dict(os.environ, NODE_NAME=NODE, GPUSHARE_PLUGIN_DIR=tmp + "/", GPUSHARE_DUMP_DIR=tmp, GPUSHARE_RETRY_SLEEP_MS="1", GSBD_ALLOW_FAKE_INVENTORY="1",
               ASAN_OPTIONS="detect_leaks=0:log_path=/tmp/httpfuzz_asan", UBSAN_OPTIONS="print_stacktrace=1:log_path=/tmp/httpfuzz_ubsan")
    env.pop("KUBECONFIG", None)
    p = subprocess.Popen([BIN, "--fake-inventory", "8", "--kube-api-url", f"http://127.0.0.1:{port}", "--timeout", "2"], env=env, stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
    runs += 1
    try:
        kubelet.register_requests.get(timeout=3); registered += 1
        ch = kubelet.channel("aliyungpushare.sock")
        for _ in range(30):
            try:
                kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]]), timeout=5); answered += 1
            except Exception:
                pass
        ch.close()
    except Exception:
        pass
    p.terminate()
    try: rc = p.wait(10)
    except subprocess.TimeoutExpired: p.kill(); rc = "hung"
    codes[rc] = codes.get(rc, 0) + 1
    kubelet.stop()
print("runs", runs, "registered", registered, "answered", answered, "exit codes", codes)
