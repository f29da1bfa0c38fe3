// alloc_load — native load generator for Allocate(): `concurrency` client threads,
// each with ONE persistent HTTP/2 connection to the plugin's unix socket (what the kubelet's device manager
// holds), issuing unary /v1beta1.DevicePlugin/Allocate calls back to back. Same output JSON as
// testing/allocate_load.py, without an interpreter's ~100 us per call on the client side.
//   alloc_load <socket> <concurrency> <total> <uuid,uuid,...>
#include <stdio.h>
#include <sys/resource.h>
#include <stdlib.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "h2.hpp"

namespace {

uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

// persistent unary client: one connection, stream ids 1, 3, 5, ...
class Client {
 public:
  bool connect(const std::string &path) {
    fd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd_ < 0) return false;
    sockaddr_un a;
    memset(&a, 0, sizeof a);
    a.sun_family = AF_UNIX;
    strncpy(a.sun_path, path.c_str(), sizeof a.sun_path - 1);
    if (::connect(fd_, (sockaddr *)&a, sizeof a) < 0) return false;
    std::string out(h2::kPreface, 24);
    out += h2::frame_bytes(h2::SETTINGS, 0, 0, nullptr, 0);
    return h2::write_full(fd_, out.data(), out.size());
  }
  ~Client() {
    if (fd_ >= 0) ::close(fd_);
  }
  // returns grpc-status (0 = OK) or -1; *resp = response message
  int call(const std::string &method, const std::string &req, std::string *resp) {
    const uint32_t sid = next_;
    next_ += 2;
    std::string block;
    h2::hpack_put_indexed(&block, 3);
    h2::hpack_put_indexed(&block, 6);
    h2::hpack_put_static_name(&block, 4, method);
    h2::hpack_put_static_name(&block, 1, "localhost");
    h2::hpack_put_static_name(&block, 31, "application/grpc");
    h2::hpack_put_header(&block, "te", "trailers");
    std::string out = h2::frame_bytes(h2::HEADERS, h2::F_END_HEADERS, sid, block.data(), block.size());
    std::string body(5, '\0');
    body[1] = (char)(req.size() >> 24);
    body[2] = (char)(req.size() >> 16);
    body[3] = (char)(req.size() >> 8);
    body[4] = (char)req.size();
    body += req;
    out += h2::frame_bytes(h2::DATA, h2::F_END_STREAM, sid, body.data(), body.size());
    if (!h2::write_full(fd_, out.data(), out.size())) return -1;
    std::string data, hdr;
    int status = -1;
    h2::Frame f;
    for (;;) {
      if (!h2::read_frame(fd_, &f)) return -1;
      if (f.type == h2::SETTINGS && !(f.flags & h2::F_ACK)) {
        const std::string ack = h2::frame_bytes(h2::SETTINGS, h2::F_ACK, 0, nullptr, 0);
        h2::write_full(fd_, ack.data(), ack.size());
      } else if (f.type == h2::PING && !(f.flags & h2::F_ACK)) {
        const std::string ack = h2::frame_bytes(h2::PING, h2::F_ACK, 0, f.payload.data(), f.payload.size());
        h2::write_full(fd_, ack.data(), ack.size());
      } else if (f.type == h2::DATA && f.stream == sid) {
        const uint8_t *p;
        size_t n;
        if (!h2::strip(f, &p, &n)) return -1;
        data.append((const char *)p, n);
        if (!f.payload.empty()) {  // replenish the connection window (streams are short-lived)
          const uint32_t inc = (uint32_t)f.payload.size();
          const uint8_t wu[4] = {(uint8_t)(inc >> 24), (uint8_t)(inc >> 16), (uint8_t)(inc >> 8), (uint8_t)inc};
          const std::string w = h2::frame_bytes(h2::WINDOW_UPDATE, 0, 0, wu, 4);
          h2::write_full(fd_, w.data(), w.size());
        }
        if (f.flags & h2::F_END_STREAM) break;
      } else if ((f.type == h2::HEADERS || f.type == h2::CONTINUATION) && f.stream == sid) {
        const uint8_t *p = (const uint8_t *)f.payload.data();
        size_t n = f.payload.size();
        if (f.type == h2::HEADERS && !h2::strip(f, &p, &n)) return -1;
        const bool end_stream = f.type == h2::HEADERS && (f.flags & h2::F_END_STREAM);
        hdr.append((const char *)p, n);
        if (f.flags & h2::F_END_HEADERS) {
          h2::Headers hs;
          if (!dec_.decode((const uint8_t *)hdr.data(), hdr.size(), &hs)) return -1;
          hdr.clear();
          for (auto &h : hs)
            if (h.first == "grpc-status") status = atoi(h.second.c_str());
        }
        if (end_stream) break;
      } else if (f.type == h2::RST_STREAM && f.stream == sid) {
        return -1;
      } else if (f.type == h2::GOAWAY) {
        return -1;
      }
    }
    if (resp && data.size() >= 5) resp->assign(data, 5, std::string::npos);
    return status;
  }

 private:
  int fd_ = -1;
  uint32_t next_ = 1;
  h2::HpackDecoder dec_;
};

std::string allocate_request(const std::vector<std::string> &ids) {  // AllocateRequest{container_requests:[{devicesIDs}]}
  std::string inner;
  for (auto &s : ids) {
    inner.push_back(0x0a);
    inner.push_back((char)s.size());
    inner += s;
  }
  std::string out;
  out.push_back(0x0a);
  size_t n = inner.size();
  while (n >= 0x80) {
    out.push_back((char)((n & 0x7F) | 0x80));
    n >>= 7;
  }
  out.push_back((char)n);
  return out + inner;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: alloc_load <socket> <concurrency> <total> <uuid,uuid,...>\n");
    return 64;
  }
  {  // one fd per connection: lift the soft limit to the hard one (a container's default soft limit can be 1024)
    rlimit rl;
    if (getrlimit(RLIMIT_NOFILE, &rl) == 0 && rl.rlim_cur < rl.rlim_max) {
      rl.rlim_cur = rl.rlim_max;
      setrlimit(RLIMIT_NOFILE, &rl);
    }
  }
  const std::string sock = argv[1];
  const int conc = atoi(argv[2]), total = atoi(argv[3]);
  std::vector<std::string> uuids;
  {
    std::stringstream ss(argv[4]);
    std::string u;
    while (std::getline(ss, u, ',')) uuids.push_back(u);
  }
  // ALLOC_LOAD_METHOD=/v1beta1.DevicePlugin/GetDevicePluginOptions measures the transport alone (no pod, no PATCH)
  const std::string method = getenv("ALLOC_LOAD_METHOD") ? getenv("ALLOC_LOAD_METHOD") : "/v1beta1.DevicePlugin/Allocate";
  std::vector<std::vector<double>> lat(conc);
  std::vector<int> errs(conc, 0), failed(conc, 0);
  std::vector<Client> clients(conc);
  for (int i = 0; i < conc; i++)
    if (!clients[i].connect(sock)) {
      fprintf(stderr, "connect failed\n");
      return 1;
    }
  std::mutex mu;
  std::condition_variable cv;
  int ready = 0;
  bool go = false;
  std::vector<std::thread> ts;
  for (int i = 0; i < conc; i++) {
    ts.emplace_back([&, i] {
      std::vector<std::string> ids;
      for (int j = 0; j < 4; j++) ids.push_back(uuids[i % uuids.size()] + "-_-" + std::to_string(j));
      const std::string req = allocate_request(ids);
      const int mine = total / conc + (i < total % conc ? 1 : 0);
      // the kubelet's connection is long-lived: the first exchange on a connection (SETTINGS, HPACK state, the
      // server's per-connection thread) is not part of an Allocate, so it is spent on a stateless RPC first
      std::string warm;
      clients[i].call("/v1beta1.DevicePlugin/GetDevicePluginOptions", "", &warm);
      {
        std::unique_lock<std::mutex> lk(mu);
        ready++;
        cv.notify_all();
        cv.wait(lk, [&] { return go; });
      }
      for (int k = 0; k < mine; k++) {
        std::string resp;
        const uint64_t t0 = now_ns();
        const int st = clients[i].call(method, req, &resp);
        lat[i].push_back((now_ns() - t0) / 1e3);
        if (st != 0) failed[i]++;
        if (resp.find("no-gpu-has") != std::string::npos) errs[i]++;
      }
    });
  }
  uint64_t t0;
  {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return ready == conc; });
    t0 = now_ns();
    go = true;
    cv.notify_all();
  }
  for (auto &t : ts) t.join();
  const double wall = (now_ns() - t0) / 1e9;
  std::vector<double> flat;
  int e = 0, fl = 0;
  for (int i = 0; i < conc; i++) {
    flat.insert(flat.end(), lat[i].begin(), lat[i].end());
    e += errs[i];
    fl += failed[i];
  }
  if (getenv("ALLOC_LOAD_DEBUG")) {
    for (int i = 0; i < conc; i++)
      for (size_t k = 0; k < lat[i].size(); k++)
        if (lat[i][k] > 2000) fprintf(stderr, "slow: thread %d call %zu %.0f us\n", i, k, lat[i][k]);
  }
  std::sort(flat.begin(), flat.end());
  auto pick = [&](double q) { return flat.empty() ? 0.0 : flat[std::min(flat.size() - 1, (size_t)(q * flat.size()))]; };
  double sum = 0;
  for (double v : flat) sum += v;
  printf("{\"concurrency\": %d, \"requests\": %zu, \"p50_us\": %.3f, \"p99_us\": %.3f, \"mean_us\": %.3f, \"req_per_s\": %.3f, "
         "\"error_responses\": %d, \"rpc_failures\": %d, \"client\": \"native\"}\n",
         conc, flat.size(), pick(0.5), pick(0.99), flat.empty() ? 0.0 : sum / flat.size(), flat.size() / wall, e, fl);
  return fl ? 2 : 0;
}
