// gsb_mock_kube — loopback stand-in for the apiserver, for the Allocate benchmark only (bench.py).
// It speaks just enough of the REST surface the plugin's Allocate path touches, and it is compiled so
// that the benchmark measures the plugin under test and not a Python HTTP server:
//   GET   /api/v1/nodes/<name>                           node object
//   GET   /api/v1/pods?fieldSelector=...                 PodList (spec.nodeName / status.phase selectors)
//   GET   /api/v1/pods?watch=true&...&resourceVersion=N  chunked stream of {"type","object"} lines
//   GET   /pods/                                         kubelet-style full PodList
//   PATCH /api/v1/namespaces/<ns>/pods/<name>            strategic merge of metadata.annotations
//   PATCH /api/v1/nodes/<name>/status                    merge of status.capacity / status.allocatable
// State is the same synthetic world testing/mock_kube.py builds;
// tests/test_native_mock.py holds the two mocks against each other request by request.
//   gsb_mock_kube --node b200-0 --pods 1024 [--mod] [--pad BYTES]   prints the port, serves until stdin closes
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <signal.h>
#include <stdio.h>
#include <sys/resource.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "json.hpp"

namespace {

struct Pod {
  std::string ns, name, node, phase;
  json::Value obj;
  std::string body;  // serialised once per change
};

struct Event {
  long rv;
  std::string type, body, node, phase;
};

struct World {
  std::mutex mu;
  std::condition_variable cv;
  long rv = 1000;
  json::Value node;
  std::string node_name;
  std::vector<Pod> pods;
  std::map<std::string, size_t> by_key;  // "<ns>/<name>"
  std::vector<Event> events;
  std::atomic<bool> closing{false};
} W;

json::Value jstr(const std::string &s) {
  json::Value v;
  v.type = json::Value::String;
  v.s = s;
  return v;
}

json::Value jobj() {
  json::Value v;
  v.type = json::Value::Object;
  return v;
}

json::Value *child(json::Value *o, const std::string &key) {  // get-or-create an object member
  for (auto &kv : o->obj)
    if (kv.first == key) return &kv.second;
  o->obj.emplace_back(key, jobj());
  return &o->obj.back().second;
}

void set_member(json::Value *o, const std::string &key, const json::Value &v) {
  for (auto &kv : o->obj)
    if (kv.first == key) {
      kv.second = v;
      return;
    }
  o->obj.emplace_back(key, v);
}

void merge_members(json::Value *dst, const json::Value *src) {  // null deletes, as a merge patch does
  if (!src || src->type != json::Value::Object) return;
  for (auto &kv : src->obj) {
    if (kv.second.type == json::Value::Null) {
      for (size_t i = 0; i < dst->obj.size(); i++)
        if (dst->obj[i].first == kv.first) {
          dst->obj.erase(dst->obj.begin() + (long)i);
          break;
        }
    } else {
      set_member(dst, kv.first, kv.second);
    }
  }
}

Pod make_pod(int i, const std::string &node, int gpu_mem, int idx, long long assume_time, size_t pad) {
  char name[32], uid[48];
  snprintf(name, sizeof name, "pod-%02d", i);
  snprintf(uid, sizeof uid, "uid-default-%05d", i);
  Pod p;
  p.ns = "default";
  p.name = name;
  p.node = node;
  p.phase = "Pending";
  json::Value o = jobj();
  set_member(&o, "kind", jstr("Pod"));
  set_member(&o, "apiVersion", jstr("v1"));
  json::Value *md = child(&o, "metadata");
  set_member(md, "name", jstr(name));
  set_member(md, "namespace", jstr("default"));
  set_member(md, "uid", jstr(uid));
  json::Value *ann = child(md, "annotations");
  set_member(ann, "ALIYUN_COM_GPU_MEM_IDX", jstr(std::to_string(idx)));
  set_member(ann, "ALIYUN_COM_GPU_MEM_ASSUME_TIME", jstr(std::to_string(assume_time)));
  set_member(ann, "ALIYUN_COM_GPU_MEM_ASSIGNED", jstr("false"));
  if (pad) set_member(ann, "example.com/padding", jstr(std::string(pad, 'x')));  // real pods carry kilobytes of managedFields
  json::Value *spec = child(&o, "spec");
  set_member(spec, "nodeName", jstr(node));
  json::Value c = jobj();
  set_member(&c, "name", jstr("c0"));
  set_member(&c, "image", jstr("busybox"));
  set_member(child(child(&c, "resources"), "limits"), "aliyun.com/gpu-mem", jstr(std::to_string(gpu_mem)));
  json::Value cs;
  cs.type = json::Value::Array;
  cs.arr.push_back(c);
  set_member(spec, "containers", cs);
  set_member(child(&o, "status"), "phase", jstr("Pending"));
  p.obj = std::move(o);
  json::dump(p.obj, &p.body);
  return p;
}

void build_world(const std::string &node, int n_pods, bool mod, size_t pad) {
  W.node_name = node;
  W.node = jobj();
  set_member(&W.node, "kind", jstr("Node"));
  set_member(&W.node, "apiVersion", jstr("v1"));
  json::Value *md = child(&W.node, "metadata");
  set_member(md, "name", jstr(node));
  child(md, "labels");
  json::Value *st = child(&W.node, "status");
  for (const char *k : {"capacity", "allocatable"}) {
    json::Value *c = child(st, k);
    set_member(c, "cpu", jstr("128"));
    set_member(c, "memory", jstr("2113929216Ki"));
    set_member(c, "aliyun.com/gpu-count", jstr("8"));
  }
  for (int i = 0; i < n_pods; i++) {
    W.pods.push_back(make_pod(i, node, 4, mod ? i % 8 : i / 8, 1700000000000000000LL + i, pad));
    W.by_key[W.pods.back().ns + "/" + W.pods.back().name] = W.pods.size() - 1;
  }
}

// caller holds W.mu
const std::string &emit(const char *type, Pod *p) {
  W.rv++;
  set_member(child(&p->obj, "metadata"), "resourceVersion", jstr(std::to_string(W.rv)));
  p->body.clear();
  json::dump(p->obj, &p->body);
  W.events.push_back({W.rv, type, p->body, p->node, p->phase});
  return p->body;  // the caller wakes the watchers once its own response is on the wire
}

// ---- HTTP/1.1
bool send_all(int fd, const std::string &s) {
  size_t off = 0;
  while (off < s.size()) {
    ssize_t n = send(fd, s.data() + off, s.size() - off, MSG_NOSIGNAL);
    if (n <= 0) return false;
    off += (size_t)n;
  }
  return true;
}

bool respond(int fd, int code, const std::string &body) {
  std::string out = "HTTP/1.1 " + std::to_string(code) + " X\r\nContent-Type: application/json\r\nContent-Length: " +
                    std::to_string(body.size()) + "\r\n\r\n";
  out += body;
  return send_all(fd, out);  // one write: headers and body in the same segment
}

bool status(int fd, int code, const std::string &message) {
  std::string b = "{\"kind\":\"Status\",\"apiVersion\":\"v1\",\"status\":\"Failure\",\"message\":";
  json::dump_string(message, &b);
  b += ",\"code\":" + std::to_string(code) + "}";
  return respond(fd, code, b);
}

std::string url_decode(const std::string &s) {
  std::string o;
  for (size_t i = 0; i < s.size(); i++) {
    if (s[i] == '%' && i + 2 < s.size() + 0 && isxdigit((unsigned char)s[i + 1]) && isxdigit((unsigned char)s[i + 2])) {
      o.push_back((char)strtol(s.substr(i + 1, 2).c_str(), nullptr, 16));
      i += 2;
    } else if (s[i] == '+') {
      o.push_back(' ');
    } else {
      o.push_back(s[i]);
    }
  }
  return o;
}

std::map<std::string, std::string> parse_query(const std::string &q) {
  std::map<std::string, std::string> m;
  size_t i = 0;
  while (i < q.size()) {
    size_t amp = q.find('&', i);
    if (amp == std::string::npos) amp = q.size();
    const std::string kv = q.substr(i, amp - i);
    const size_t eq = kv.find('=');
    if (eq == std::string::npos) m[url_decode(kv)] = "";
    else m[url_decode(kv.substr(0, eq))] = url_decode(kv.substr(eq + 1));
    i = amp + 1;
  }
  return m;
}

std::map<std::string, std::string> parse_selector(const std::string &sel) {
  std::map<std::string, std::string> m;
  size_t i = 0;
  while (i < sel.size()) {
    size_t comma = sel.find(',', i);
    if (comma == std::string::npos) comma = sel.size();
    const std::string kv = sel.substr(i, comma - i);
    const size_t eq = kv.find('=');
    if (eq != std::string::npos) m[kv.substr(0, eq)] = kv.substr(eq + 1);
    i = comma + 1;
  }
  return m;
}

bool matches(const std::map<std::string, std::string> &sel, const std::string &node, const std::string &phase) {
  auto n = sel.find("spec.nodeName");
  auto p = sel.find("status.phase");
  return (n == sel.end() || n->second == node) && (p == sel.end() || p->second == phase);
}

std::vector<std::string> split_path(const std::string &path) {
  std::vector<std::string> parts;
  size_t i = 0;
  while (i < path.size()) {
    size_t s = path.find('/', i);
    if (s == std::string::npos) s = path.size();
    if (s > i) parts.push_back(path.substr(i, s - i));
    i = s + 1;
  }
  return parts;
}

void watch(int fd, const std::map<std::string, std::string> &sel, long since) {
  if (!send_all(fd, "HTTP/1.1 200 OK\r\nContent-Type: application/json\r\nTransfer-Encoding: chunked\r\n\r\n")) return;
  size_t cursor = 0;
  while (true) {
    std::vector<Event> batch;
    {
      std::unique_lock<std::mutex> lk(W.mu);
      while (cursor >= W.events.size() && !W.closing) {
        W.cv.wait_for(lk, std::chrono::milliseconds(200));
        char probe;  // a peer that has gone away ends the stream (recv returns 0 / error without blocking)
        const ssize_t r = recv(fd, &probe, 1, MSG_PEEK | MSG_DONTWAIT);
        if (r == 0 || (r < 0 && errno != EAGAIN && errno != EWOULDBLOCK)) return;
      }
      if (W.closing) break;
      batch.assign(W.events.begin() + (long)cursor, W.events.end());
      cursor = W.events.size();
    }
    std::string out;
    for (auto &e : batch) {
      if (e.rv <= since) continue;
      const std::string type = matches(sel, e.node, e.phase) ? e.type : "DELETED";
      const std::string line = "{\"type\":\"" + type + "\",\"object\":" + e.body + "}\n";
      char head[32];
      snprintf(head, sizeof head, "%zx\r\n", line.size());
      out += head;
      out += line;
      out += "\r\n";
    }
    if (!out.empty() && !send_all(fd, out)) return;
  }
  send_all(fd, "0\r\n\r\n");
}

// returns false when the connection should close
bool handle(int fd, const std::string &method, const std::string &target, const std::string &body) {
  const size_t qm = target.find('?');
  const std::string path = target.substr(0, qm), query = qm == std::string::npos ? "" : target.substr(qm + 1);
  const std::vector<std::string> parts = split_path(path);
  const bool api = parts.size() >= 3 && parts[0] == "api" && parts[1] == "v1";
  if (method == "GET") {
    if (path == "/pods" || path == "/pods/") {
      std::string out = "{\"kind\":\"PodList\",\"apiVersion\":\"v1\",\"items\":[";
      {
        std::lock_guard<std::mutex> lk(W.mu);
        for (size_t i = 0; i < W.pods.size(); i++) {
          if (i) out.push_back(',');
          out += W.pods[i].body;
        }
      }
      return respond(fd, 200, out + "]}");
    }
    if (api && parts[2] == "nodes" && parts.size() == 4) {
      std::string out;
      {
        std::lock_guard<std::mutex> lk(W.mu);
        if (parts[3] != W.node_name) return status(fd, 404, "nodes \"" + parts[3] + "\" not found");
        json::dump(W.node, &out);
      }
      return respond(fd, 200, out);
    }
    if (api && parts[2] == "nodes" && parts.size() == 3) {
      std::string out = "{\"kind\":\"NodeList\",\"apiVersion\":\"v1\",\"items\":[";
      {
        std::lock_guard<std::mutex> lk(W.mu);
        json::dump(W.node, &out);
      }
      return respond(fd, 200, out + "]}");
    }
    if (api && parts[2] == "pods" && parts.size() == 3) {
      const auto q = parse_query(query);
      const auto sel = parse_selector(q.count("fieldSelector") ? q.at("fieldSelector") : "");
      const std::string w = q.count("watch") ? q.at("watch") : "";
      if (w == "true" || w == "1") {
        watch(fd, sel, q.count("resourceVersion") ? atol(q.at("resourceVersion").c_str()) : 0);
        return false;
      }
      std::string out;
      {
        std::lock_guard<std::mutex> lk(W.mu);
        out = "{\"kind\":\"PodList\",\"apiVersion\":\"v1\",\"metadata\":{\"resourceVersion\":\"" + std::to_string(W.rv) +
              "\"},\"items\":[";
        bool first = true;
        for (auto &p : W.pods) {
          if (!matches(sel, p.node, p.phase)) continue;
          if (!first) out.push_back(',');
          first = false;
          out += p.body;
        }
      }
      return respond(fd, 200, out + "]}");
    }
    return status(fd, 404, "not found");
  }
  if (method == "PATCH") {
    json::Value patch;
    if (!json::parse(body, &patch)) return status(fd, 400, "invalid JSON patch");
    if (api && parts[2] == "nodes" && parts.size() == 5 && parts[4] == "status") {
      std::string out;
      {
        std::lock_guard<std::mutex> lk(W.mu);
        if (parts[3] != W.node_name) return status(fd, 404, "nodes \"" + parts[3] + "\" not found");
        json::Value *st = child(&W.node, "status");
        for (const char *k : {"capacity", "allocatable"}) merge_members(child(st, k), patch.path({"status", k}));
        json::dump(W.node, &out);
      }
      return respond(fd, 200, out);
    }
    if (api && parts[2] == "namespaces" && parts.size() == 6 && parts[4] == "pods") {
      std::string out;
      {
        std::lock_guard<std::mutex> lk(W.mu);
        auto it = W.by_key.find(parts[3] + "/" + parts[5]);
        if (it == W.by_key.end()) return status(fd, 404, "pods \"" + parts[5] + "\" not found");
        Pod *p = &W.pods[it->second];
        merge_members(child(child(&p->obj, "metadata"), "annotations"), patch.path({"metadata", "annotations"}));
        out = emit("MODIFIED", p);
      }
      const bool ok = respond(fd, 200, out);
      W.cv.notify_all();  // watch fan-out is asynchronous to the writer's response, as in the apiserver's watch cache
      return ok;
    }
    return status(fd, 404, "not found");
  }
  return status(fd, 405, "method not allowed");
}

void serve_conn(int fd) {
  const int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
  std::string buf;
  char tmp[16384];
  while (!W.closing) {
    size_t hdr_end;
    while ((hdr_end = buf.find("\r\n\r\n")) == std::string::npos) {
      const ssize_t n = recv(fd, tmp, sizeof tmp, 0);
      if (n <= 0) goto out;
      buf.append(tmp, (size_t)n);
      if (buf.size() > (1u << 20)) goto out;
    }
    {
      const std::string head = buf.substr(0, hdr_end);
      const size_t sp1 = head.find(' '), sp2 = head.find(' ', sp1 + 1);
      if (sp1 == std::string::npos || sp2 == std::string::npos) goto out;
      const std::string method = head.substr(0, sp1), target = head.substr(sp1 + 1, sp2 - sp1 - 1);
      size_t clen = 0;
      bool close_after = false;
      size_t line = head.find("\r\n");
      while (line != std::string::npos && line + 2 < head.size()) {
        size_t next = head.find("\r\n", line + 2);
        std::string h = head.substr(line + 2, (next == std::string::npos ? head.size() : next) - line - 2);
        for (auto &c : h) {
          if (c == ':') break;
          c = (char)tolower((unsigned char)c);
        }
        if (h.rfind("content-length:", 0) == 0) clen = (size_t)atol(h.c_str() + 15);
        if (h.rfind("connection:", 0) == 0 && h.find("close") != std::string::npos) close_after = true;
        line = next;
      }
      if (clen > (64u << 20)) goto out;
      while (buf.size() < hdr_end + 4 + clen) {
        const ssize_t n = recv(fd, tmp, sizeof tmp, 0);
        if (n <= 0) goto out;
        buf.append(tmp, (size_t)n);
      }
      const std::string body = buf.substr(hdr_end + 4, clen);
      buf.erase(0, hdr_end + 4 + clen);
      if (!handle(fd, method, target, body) || close_after) goto out;
    }
  }
out:
  close(fd);
}

}  // namespace

int main(int argc, char **argv) {
  std::string node = "b200-0";
  int n_pods = 64;
  bool mod = false;
  size_t pad = 0;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "--node" && i + 1 < argc) node = argv[++i];
    else if (a == "--pods" && i + 1 < argc) n_pods = atoi(argv[++i]);
    else if (a == "--pad" && i + 1 < argc) pad = (size_t)atol(argv[++i]);
    else if (a == "--mod") mod = true;
    else {
      fprintf(stderr, "usage: gsb_mock_kube [--node NAME] [--pods N] [--mod] [--pad BYTES]\n");
      return 64;
    }
  }
  {  // one fd per connection: lift the soft limit to the hard one (a container's default soft limit can be 1024)
    rlimit rl;
    if (getrlimit(RLIMIT_NOFILE, &rl) == 0 && rl.rlim_cur < rl.rlim_max) {
      rl.rlim_cur = rl.rlim_max;
      setrlimit(RLIMIT_NOFILE, &rl);
    }
  }
  signal(SIGPIPE, SIG_IGN);
  build_world(node, n_pods, mod, pad);
  const int lfd = socket(AF_INET, SOCK_STREAM, 0);
  const int one = 1;
  setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
  addr.sin_port = 0;
  if (bind(lfd, (sockaddr *)&addr, sizeof addr) != 0 || listen(lfd, 1024) != 0) {
    perror("bind/listen");
    return 1;
  }
  socklen_t alen = sizeof addr;
  getsockname(lfd, (sockaddr *)&addr, &alen);
  printf("%d\n", ntohs(addr.sin_port));
  fflush(stdout);
  std::thread acceptor([lfd] {
    while (true) {
      const int fd = accept(lfd, nullptr, nullptr);
      if (fd < 0) {
        if (W.closing) return;
        continue;
      }
      std::thread(serve_conn, fd).detach();
    }
  });
  while (getchar() != EOF) {
  }
  W.closing = true;
  W.cv.notify_all();
  shutdown(lfd, SHUT_RDWR);
  close(lfd);
  acceptor.join();
  _exit(0);  // connection threads are detached; the process is the unit of clean-up
}
