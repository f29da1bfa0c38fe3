"""Random HTTP/2 frame sequences (valid-ish HEADERS/DATA mixed with damaged frames, SETTINGS, WINDOW_UPDATE, RST_STREAM,
CONTINUATION) against the h2 self-test server, normally its ASan/UBSan build:
    python tools/fuzz_h2.py build/san/h2.asan <seconds> <seed>   (tools/sanitize.sh builds build/san/h2.asan)
The server must still answer a normal call afterwards; sanitizer reports land in /tmp/h2fuzz_asan.* / _ubsan.*"""
import os, random, socket, subprocess, sys, time
BIN = sys.argv[1]; secs = float(sys.argv[2]); seed = int(sys.argv[3])
rnd = random.Random(seed)
sock_path = f"/tmp/h2fuzz-{os.getpid()}.sock"
env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:log_path=/tmp/h2fuzz_asan", UBSAN_OPTIONS="print_stacktrace=1:log_path=/tmp/h2fuzz_ubsan")
p = subprocess.Popen([BIN, "serve", sock_path], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env)
p.stdout.readline()
def frame(t, fl, sid, pl=b""): return len(pl).to_bytes(3, "big") + bytes([t, fl]) + (sid & 0x7FFFFFFF).to_bytes(4, "big") + pl
paths = [b"/test.Echo/Unary", b"/test.Echo/Stream", b"/test.Echo/Fail", b"/nope", b"/test.Echo/Forever"]
def headers(path): return bytes([0x83, 0x86, 0x04, len(path)]) + path + bytes([0x0f, 0x10, 16]) + b"application/grpc"
end = time.time() + secs; conns = 0
while time.time() < end:
    if p.poll() is not None:
        print("SERVER DIED", p.returncode, "seed", seed, "conns", conns); sys.exit(1)
    s = socket.socket(socket.AF_UNIX); s.settimeout(0.05)
    try:
        s.connect(sock_path)
        if rnd.random() < 0.9: s.sendall(b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n")
        sid = 1
        for _ in range(rnd.randint(1, 40)):
            r = rnd.random()
            if r < 0.35:
                fl = rnd.choice([0x4, 0x5, 0x0, 0x4 | 0x8, 0x4 | 0x20, 0x24])
                pl = headers(rnd.choice(paths))
                if fl & 0x8: pl = bytes([rnd.randint(0, 9)]) + pl + b"\0" * rnd.randint(0, 9)
                if fl & 0x20: pl = bytes(rnd.getrandbits(8) for _ in range(5)) + pl
                s.sendall(frame(1, fl, sid, pl)); 
                if rnd.random() < 0.7: sid += 2
            elif r < 0.6:
                n = rnd.choice([0, 1, 5, 9, 100, 20000])
                body = b"\0" + rnd.choice([n, n + 1, 0, 1 << 30]).to_bytes(4, "big") + bytes(rnd.getrandbits(8) for _ in range(min(n, 300))) + b"x" * max(0, n - 300)
                s.sendall(frame(0, rnd.choice([0, 1, 8, 9]), rnd.choice([sid, sid - 2, 1, 0, 99]), body))
            elif r < 0.7:
                s.sendall(frame(rnd.randint(0, 12), rnd.getrandbits(8), rnd.choice([0, 1, sid, 2, 1 << 31]), bytes(rnd.getrandbits(8) for _ in range(rnd.choice([0, 1, 4, 5, 6, 8, 9, 30])))))
            elif r < 0.8:
                s.sendall(frame(4, rnd.choice([0, 1]), 0, b"".join(rnd.choice([1, 2, 3, 4, 5, 6, 9]).to_bytes(2, "big") + rnd.choice([0, 1, 100, 16384, 1 << 24, (1 << 31) - 1, (1 << 32) - 1]).to_bytes(4, "big") for _ in range(rnd.randint(0, 4)))))
            elif r < 0.88:
                s.sendall(frame(8, 0, rnd.choice([0, 1, sid]), rnd.choice([0, 1, 1000, (1 << 31) - 1]).to_bytes(4, "big")))
            elif r < 0.94:
                s.sendall(frame(3, 0, rnd.choice([1, sid, sid - 2]), rnd.randint(0, 13).to_bytes(4, "big")))
            else:
                s.sendall(frame(9, rnd.choice([0, 4]), rnd.choice([1, sid]), bytes(rnd.getrandbits(8) for _ in range(rnd.randint(0, 20)))))
            if rnd.random() < 0.2:
                try: s.recv(65536)
                except OSError: pass
    except OSError:
        pass
    finally:
        s.close(); conns += 1
# still serving?
o = subprocess.run([BIN, "call", sock_path, "/test.Echo/Unary", "6869"], capture_output=True, text=True)
print("conns", conns, "final call:", o.stdout.strip(), "alive" if p.poll() is None else f"dead {p.returncode}")
p.stdin.close(); 
try: p.wait(5)
except Exception: p.kill()
os.unlink(sock_path) if os.path.exists(sock_path) else None
