"""The deployment path only CPU tests can cover: HTTPS to the apiserver with CA verification, name/IP
check, bearer token, client certificates and the kubeconfig forms clusters really use (file paths and
base64 *-data fields) — for the native daemon's OpenSSL client and for the Python kube client."""
import base64
import os
import subprocess

import pytest

from gpushare_device_plugin_b200.nvidia import kubeclient
from gpushare_device_plugin_b200.testing.fake_kubelet import FakeKubelet
from gpushare_device_plugin_b200.testing.mock_kube import MockKube, config4_pods, make_node
from oracle import wire_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GSBD = os.environ.get("GSBD_BINARY") or os.path.join(ROOT, "gpushare_device_plugin_b200", "gsbd")
NODE = "b200-0"


def sh(*a, cwd):
    subprocess.run(a, cwd=cwd, check=True, capture_output=True)


@pytest.fixture(scope="module")
def pki(tmp_path_factory):
    d = tmp_path_factory.mktemp("pki")
    sh("openssl", "req", "-x509", "-newkey", "rsa:2048", "-nodes", "-keyout", "ca.key", "-out", "ca.crt", "-subj", "/CN=test-ca",
       "-days", "2", cwd=d)
    for name, san in (("server", "IP:127.0.0.1"), ("wrongname", "DNS:not-this-host"), ("client", "DNS:client")):
        sh("openssl", "req", "-newkey", "rsa:2048", "-nodes", "-keyout", f"{name}.key", "-out", f"{name}.csr", "-subj", f"/CN={name}", cwd=d)
        (d / f"{name}.ext").write_text(f"subjectAltName={san}\n")
        sh("openssl", "x509", "-req", "-in", f"{name}.csr", "-CA", "ca.crt", "-CAkey", "ca.key", "-CAcreateserial", "-out",
           f"{name}.crt", "-days", "2", "-extfile", f"{name}.ext", cwd=d)
    sh("openssl", "req", "-x509", "-newkey", "rsa:2048", "-nodes", "-keyout", "other.key", "-out", "other-ca.crt", "-subj",
       "/CN=other-ca", "-days", "2", cwd=d)
    return d


def kubeconfig(path, server, cluster_lines, user_lines):
    path.write_text("apiVersion: v1\nkind: Config\ncurrent-context: c\nclusters:\n- name: k\n  cluster:\n"
                    f"    server: {server}\n" + "".join(f"    {l}\n" for l in cluster_lines) +
                    "contexts:\n- name: c\n  context:\n    cluster: k\n    user: u\nusers:\n- name: u\n  user:\n" +
                    "".join(f"    {l}\n" for l in user_lines))
    return str(path)


def run_gsbd(tmp_path, kc, expect_register=True):
    kubelet = FakeKubelet(str(tmp_path))
    env = dict(os.environ, NODE_NAME=NODE, KUBECONFIG=kc, GPUSHARE_PLUGIN_DIR=str(tmp_path) + "/", GPUSHARE_RETRY_SLEEP_MS="1",
               GSBD_ALLOW_FAKE_INVENTORY="1")
    log = open(tmp_path / "gsbd.log", "w")
    p = subprocess.Popen([GSBD, "--v=5", "--fake-inventory", "8"], env=env, stderr=log, stdout=log)
    try:
        if expect_register:
            kubelet.register_requests.get(timeout=20)
            ch = kubelet.channel("aliyungpushare.sock")
            envs = wo.unmarshal_AllocateResponse(kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))
            ch.close()
            return envs
        return p.wait(timeout=20), open(tmp_path / "gsbd.log").read()
    finally:
        if p.poll() is None:
            p.terminate()
            p.wait(timeout=10)
        log.close()
        kubelet.stop()


def b64file(p):
    return base64.b64encode(open(p, "rb").read()).decode()


@pytest.mark.skipif(not os.access(GSBD, os.X_OK), reason="gsbd not built")
class TestNativeDaemonTls:
    def test_ca_file_and_token(self, pki, tmp_path):
        kube = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "server.crt"), str(pki / "server.key")))
        try:
            kc = kubeconfig(tmp_path / "kc", kube.url, [f"certificate-authority: {pki / 'ca.crt'}"], ["token: s3cr3t-token"])
            envs = run_gsbd(tmp_path, kc)
            assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
            assert "Bearer s3cr3t-token" in kube.auth_headers
            assert kube.pod("pod-00")["metadata"]["annotations"]["ALIYUN_COM_GPU_MEM_ASSIGNED"] == "true"
        finally:
            kube.close()

    def test_embedded_data_fields_and_client_certificate(self, pki, tmp_path):
        kube = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "server.crt"), str(pki / "server.key")),
                        client_ca=str(pki / "ca.crt"))
        try:
            kc = kubeconfig(tmp_path / "kc", kube.url, [f"certificate-authority-data: {b64file(pki / 'ca.crt')}"],
                            [f"client-certificate-data: {b64file(pki / 'client.crt')}",
                             f"client-key-data: {b64file(pki / 'client.key')}"])
            assert run_gsbd(tmp_path, kc)[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
            # same server, no client certificate: the handshake is refused, the daemon exits 1 (gpumanager.go:73)
            kc2 = kubeconfig(tmp_path / "kc2", kube.url, [f"certificate-authority: {pki / 'ca.crt'}"], ["token: t"])
            rc, log = run_gsbd(tmp_path, kc2, expect_register=False)
            assert rc == 1 and "Failed to get device plugin" in log
        finally:
            kube.close()

    def test_wrong_ca_and_wrong_name_are_refused_unless_insecure(self, pki, tmp_path):
        kube = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "server.crt"), str(pki / "server.key")))
        try:
            kc = kubeconfig(tmp_path / "kc", kube.url, [f"certificate-authority: {pki / 'other-ca.crt'}"], ["token: t"])
            rc, log = run_gsbd(tmp_path, kc, expect_register=False)
            assert rc == 1 and "x509" in log
            kc = kubeconfig(tmp_path / "kc3", kube.url, ["insecure-skip-tls-verify: true"], ["token: t"])
            assert run_gsbd(tmp_path, kc)[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
        finally:
            kube.close()
        kube = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "wrongname.crt"), str(pki / "wrongname.key")))
        try:  # right CA, but the certificate is for another host
            kc = kubeconfig(tmp_path / "kc4", kube.url, [f"certificate-authority: {pki / 'ca.crt'}"], ["token: t"])
            rc, log = run_gsbd(tmp_path, kc, expect_register=False)
            assert rc == 1 and "x509" in log
        finally:
            kube.close()


@pytest.mark.skipif(not os.access(GSBD, os.X_OK), reason="gsbd not built")
def test_in_cluster_config_and_rotated_service_account_token(pki, tmp_path):
    """rest.InClusterConfig (podmanager.go:32-40): KUBERNETES_SERVICE_HOST/PORT + the mounted token and ca.crt. The
    kubelet rewrites a projected token before it expires; a 401 makes the daemon read the file again, once."""
    kube = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "server.crt"), str(pki / "server.key")))
    kube.required_token = "token-1"
    sa = tmp_path / "sa"
    sa.mkdir()
    (sa / "token").write_text("token-1\n")
    (sa / "ca.crt").write_bytes((pki / "ca.crt").read_bytes())
    kubelet = FakeKubelet(str(tmp_path))
    env = dict(os.environ, NODE_NAME=NODE, KUBERNETES_SERVICE_HOST="127.0.0.1", KUBERNETES_SERVICE_PORT=str(kube.port),
               GSBD_SERVICEACCOUNT_DIR=str(sa), GPUSHARE_PLUGIN_DIR=str(tmp_path) + "/", GPUSHARE_RETRY_SLEEP_MS="1",
               GSBD_ALLOW_FAKE_INVENTORY="1")
    env.pop("KUBECONFIG", None)
    log = open(tmp_path / "gsbd.log", "w")
    p = subprocess.Popen([GSBD, "--v=5", "--fake-inventory", "8", "--pod-informer=false"], env=env, stderr=log, stdout=log)
    try:
        kubelet.register_requests.get(timeout=20)
        ch = kubelet.channel("aliyungpushare.sock")
        req = wo.marshal_AllocateRequest([["a", "b", "c", "d"]])
        assert wo.unmarshal_AllocateResponse(kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
        # rotation: the apiserver stops honouring the old token, the new one is on disk
        (sa / "token").write_text("token-2\n")
        kube.required_token = "token-2"
        for _ in range(3):
            assert wo.unmarshal_AllocateResponse(kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
        assert kube.auth_headers.count("Bearer token-1") >= 3 and "Bearer token-2" in kube.auth_headers
        # a token that is simply wrong (file unchanged) is not retried for ever: the request is refused
        kube.required_token = "token-3"
        assert wo.unmarshal_AllocateResponse(kubelet.allocate(ch, req))[0]["ALIYUN_COM_GPU_MEM_IDX"] == "-1"
        ch.close()
    finally:
        p.terminate()
        p.wait(timeout=10)
        log.close()
        kubelet.stop()
        kube.close()


def test_python_kube_client_tls_forms(pki, tmp_path, monkeypatch):
    kube = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "server.crt"), str(pki / "server.key")),
                    client_ca=str(pki / "ca.crt"))
    try:
        kc = kubeconfig(tmp_path / "kc", kube.url, [f"certificate-authority-data: {b64file(pki / 'ca.crt')}"],
                        [f"client-certificate-data: {b64file(pki / 'client.crt')}", f"client-key-data: {b64file(pki / 'client.key')}",
                         "token: tok"])
        monkeypatch.setenv("KUBECONFIG", kc)
        cs = kubeclient.from_environment()
        assert cs.get_node(NODE)["metadata"]["name"] == NODE and "Bearer tok" in kube.auth_headers
        kc = kubeconfig(tmp_path / "kc2", kube.url, [f"certificate-authority: {pki / 'other-ca.crt'}"],
                        [f"client-certificate: {pki / 'client.crt'}", f"client-key: {pki / 'client.key'}"])
        monkeypatch.setenv("KUBECONFIG", kc)
        with pytest.raises(Exception):
            kubeclient.from_environment().get_node(NODE)
    finally:
        kube.close()


def test_python_kube_client_in_cluster_and_rotated_token(pki, tmp_path, monkeypatch):
    kube = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "server.crt"), str(pki / "server.key")))
    kube.required_token = "token-1"
    try:
        sa = tmp_path / "sa"
        sa.mkdir()
        (sa / "token").write_text("token-1\n")
        (sa / "ca.crt").write_bytes((pki / "ca.crt").read_bytes())
        monkeypatch.delenv("KUBECONFIG", raising=False)
        monkeypatch.setenv("KUBERNETES_SERVICE_HOST", "127.0.0.1")
        monkeypatch.setenv("KUBERNETES_SERVICE_PORT", str(kube.port))
        monkeypatch.setattr(kubeclient, "SA_DIR", str(sa))
        cs = kubeclient.from_environment()
        assert cs.get_node(NODE)["metadata"]["name"] == NODE
        (sa / "token").write_text("token-2\n")
        kube.required_token = "token-2"
        assert cs.get_node(NODE)["metadata"]["name"] == NODE and cs.token == "token-2"
        kube.required_token = "token-3"  # simply wrong, nothing new on disk: the error surfaces
        with pytest.raises(kubeclient.ApiError) as e:
            cs.get_node(NODE)
        assert e.value.status == 401
    finally:
        kube.close()


@pytest.mark.skipif(not os.access(GSBD, os.X_OK), reason="gsbd not built")
def test_concurrent_allocates_and_watch_over_tls(pki, tmp_path):
    """The HTTPS client under concurrency: pooled TLS connections for the PATCHes, a long-lived TLS watch stream, pod
    churn feeding it. (tools/sanitize.sh runs this under TSan and ASan: one SSL_CTX shared by every connection.)"""
    import threading
    import time

    from gpushare_device_plugin_b200.testing.mock_kube import make_pod
    kube = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "server.crt"), str(pki / "server.key")))
    kc = kubeconfig(tmp_path / "kc", kube.url, [f"certificate-authority: {pki / 'ca.crt'}"], ["token: t"])
    kubelet = FakeKubelet(str(tmp_path))
    env = dict(os.environ, NODE_NAME=NODE, KUBECONFIG=kc, GPUSHARE_PLUGIN_DIR=str(tmp_path) + "/", GPUSHARE_RETRY_SLEEP_MS="1",
               GSBD_ALLOW_FAKE_INVENTORY="1")
    log = open(tmp_path / "gsbd.log", "w")
    p = subprocess.Popen([GSBD, "--v=5", "--fake-inventory", "8"], env=env, stderr=log, stdout=log)
    try:
        kubelet.register_requests.get(timeout=20)
        stop, results, errors = threading.Event(), [], []

        def allocator():
            ch = kubelet.channel("aliyungpushare.sock")
            try:
                while not stop.is_set():
                    results.append(wo.unmarshal_AllocateResponse(
                        kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b"]])))[0]["ALIYUN_COM_GPU_MEM_IDX"])
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))
            finally:
                ch.close()

        def churn():
            i = 1000
            while not stop.is_set():
                kube.add_pod(make_pod(i, NODE, gpu_mem=2, idx=i % 8, assume_time=1_800_000_000_000_000_000 + i))
                i += 1
                time.sleep(0.005)
        ts = [threading.Thread(target=allocator) for _ in range(6)] + [threading.Thread(target=churn)]
        [t.start() for t in ts]
        time.sleep(3)
        stop.set()
        [t.join(30) for t in ts]
        applied = list(kube.patched_ok)
        assert not errors and len(applied) > 100 and len(applied) == len(set(applied))
        assert len([r for r in results if r != "-1"]) == len(applied) and kube.watches_served >= 1
        p.terminate()
        assert p.wait(timeout=15) == 0
    finally:
        if p.poll() is None:
            p.kill()
        log.close()
        kubelet.stop()
        kube.close()


@pytest.mark.skipif(not os.access(GSBD, os.X_OK), reason="gsbd not built")
def test_kubelet_pods_client_presents_the_client_certificate(pki, tmp_path):
    """--query-kubelet with --client-cert/--client-key (cmd/nvidia/main.go:40-46: TLSClientConfig{CertFile, KeyFile},
    server verification off): the kubelet /pods/ endpoint demands a client certificate; the pending pods must come
    from it at the first try, not from the apiserver fallback after nine refused handshakes."""
    import time
    api = MockKube(make_node(NODE), config4_pods(NODE))
    kubelet_api = MockKube(make_node(NODE), config4_pods(NODE), tls=(str(pki / "server.crt"), str(pki / "server.key")),
                           client_ca=str(pki / "ca.crt"))
    kubelet = FakeKubelet(str(tmp_path))
    env = dict(os.environ, NODE_NAME=NODE, GPUSHARE_PLUGIN_DIR=str(tmp_path) + "/", GPUSHARE_RETRY_SLEEP_MS="1",
               GSBD_ALLOW_FAKE_INVENTORY="1")
    env.pop("KUBECONFIG", None)
    log = open(tmp_path / "gsbd.log", "w")
    p = subprocess.Popen([GSBD, "--v=5", "--fake-inventory", "8", "--kube-api-url", api.url, "--query-kubelet",
                          "--kubelet-address", "127.0.0.1", "--kubelet-port", str(kubelet_api.port),
                          "--client-cert", str(pki / "client.crt"), "--client-key", str(pki / "client.key")],
                         env=env, stderr=log, stdout=log)
    try:
        kubelet.register_requests.get(timeout=20)
        ch = kubelet.channel("aliyungpushare.sock")
        t0 = time.monotonic()
        envs = wo.unmarshal_AllocateResponse(kubelet.allocate(ch, wo.marshal_AllocateRequest([["a", "b", "c", "d"]])))
        took = time.monotonic() - t0
        ch.close()
        assert envs[0]["ALIYUN_COM_GPU_MEM_IDX"] == "0"
        assert any(r[:2] == ("GET", "/pods/") for r in kubelet_api.requests)        # served by the kubelet, over mTLS
        assert not any(r[0] == "GET" and r[1].startswith("/api/v1/pods?") for r in api.requests)  # no apiserver fallback
        assert took < 0.8  # nine failed tries would have cost 8 x 100 ms
    finally:
        if p.poll() is None:
            p.terminate()
            p.wait(timeout=10)
        log.close()
        kubelet.stop()
        api.close()
        kubelet_api.close()
