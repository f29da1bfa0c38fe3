// h2.hpp — the slice of HTTP/2 (RFC 7540) + HPACK (RFC 7541) that gRPC over a unix socket needs.
//
// Written for the native front end of the device plugin (gsbd): the kubelet's grpc-go client talks
// cleartext HTTP/2 with prior knowledge to /var/lib/kubelet/device-plugins/aliyungpushare.sock
// (reference: pkg/gpu/nvidia/server.go:106-134 uses grpc.NewServer() with default options), and the
// plugin itself is an HTTP/2 client of kubelet.sock for Registration/Register (server.go:150-169).
// Scope: SETTINGS, HEADERS/CONTINUATION, DATA with both flow-control windows, WINDOW_UPDATE, PING,
// RST_STREAM, GOAWAY; HPACK decoding with the dynamic table and Huffman strings; encoding uses
// literals without indexing (always legal). No TLS, no push, no priorities (ignored as allowed).
#pragma once

#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace h2 {

// ------------------------------------------------------------------------------------ HPACK

struct HuffSym {
  uint32_t code;
  uint8_t bits;
};
static const HuffSym kHuff[257] = {
#include "hpack_huffman_table.inc"
};

class HuffmanDecoder {
 public:
  HuffmanDecoder() {
    nodes_.push_back({{-1, -1}, -1});
    for (int sym = 0; sym < 257; sym++) {
      int n = 0;
      for (int b = kHuff[sym].bits - 1; b >= 0; b--) {
        const int bit = (kHuff[sym].code >> b) & 1;
        if (nodes_[n].child[bit] < 0) {
          nodes_[n].child[bit] = (int)nodes_.size();
          nodes_.push_back({{-1, -1}, -1});
        }
        n = nodes_[n].child[bit];
      }
      nodes_[n].sym = sym;
    }
  }
  bool decode(const uint8_t *p, size_t len, std::string *out) const {
    int n = 0, depth = 0;
    bool all_ones = true;
    for (size_t i = 0; i < len; i++) {
      for (int b = 7; b >= 0; b--) {
        const int bit = (p[i] >> b) & 1;
        n = nodes_[n].child[bit];
        if (n < 0) return false;
        depth++;
        all_ones = all_ones && bit;
        if (nodes_[n].sym >= 0) {
          if (nodes_[n].sym == 256) return false;  // EOS inside a string is an error
          out->push_back((char)nodes_[n].sym);
          n = 0;
          depth = 0;
          all_ones = true;
        }
      }
    }
    return depth < 8 && all_ones;  // padding = most significant bits of EOS
  }

 private:
  struct Node {
    int child[2];
    int sym;
  };
  std::vector<Node> nodes_;
};

inline const HuffmanDecoder &huffman() {
  static const HuffmanDecoder d;
  return d;
}

typedef std::vector<std::pair<std::string, std::string>> Headers;

static const char *const kStaticTable[62][2] = {
    {"", ""}, {":authority", ""}, {":method", "GET"}, {":method", "POST"}, {":path", "/"}, {":path", "/index.html"},
    {":scheme", "http"}, {":scheme", "https"}, {":status", "200"}, {":status", "204"}, {":status", "206"},
    {":status", "304"}, {":status", "400"}, {":status", "404"}, {":status", "500"}, {"accept-charset", ""},
    {"accept-encoding", "gzip, deflate"}, {"accept-language", ""}, {"accept-ranges", ""}, {"accept", ""},
    {"access-control-allow-origin", ""}, {"age", ""}, {"allow", ""}, {"authorization", ""}, {"cache-control", ""},
    {"content-disposition", ""}, {"content-encoding", ""}, {"content-language", ""}, {"content-length", ""},
    {"content-location", ""}, {"content-range", ""}, {"content-type", ""}, {"cookie", ""}, {"date", ""}, {"etag", ""},
    {"expect", ""}, {"expires", ""}, {"from", ""}, {"host", ""}, {"if-match", ""}, {"if-modified-since", ""},
    {"if-none-match", ""}, {"if-range", ""}, {"if-unmodified-since", ""}, {"last-modified", ""}, {"link", ""},
    {"location", ""}, {"max-forwards", ""}, {"proxy-authenticate", ""}, {"proxy-authorization", ""}, {"range", ""},
    {"referer", ""}, {"refresh", ""}, {"retry-after", ""}, {"server", ""}, {"set-cookie", ""},
    {"strict-transport-security", ""}, {"transfer-encoding", ""}, {"user-agent", ""}, {"vary", ""}, {"via", ""},
    {"www-authenticate", ""}};

class HpackDecoder {
 public:
  bool decode(const uint8_t *p, size_t len, Headers *out) {
    const uint8_t *end = p + len;
    while (p < end) {
      const uint8_t b = *p;
      if (b & 0x80) {  // indexed header field
        uint64_t idx;
        if (!integer(&p, end, 7, &idx)) return false;
        std::string n, v;
        if (!lookup(idx, &n, &v)) return false;
        out->emplace_back(std::move(n), std::move(v));
      } else if ((b & 0xC0) == 0x40) {  // literal with incremental indexing
        std::string n, v;
        if (!literal(&p, end, 6, &n, &v)) return false;
        insert(n, v);
        out->emplace_back(std::move(n), std::move(v));
      } else if ((b & 0xE0) == 0x20) {  // dynamic table size update
        uint64_t sz;
        if (!integer(&p, end, 5, &sz) || sz > settings_max_) return false;
        max_size_ = (size_t)sz;
        evict();
      } else {  // literal without indexing / never indexed
        std::string n, v;
        if (!literal(&p, end, 4, &n, &v)) return false;
        out->emplace_back(std::move(n), std::move(v));
      }
    }
    return true;
  }

 private:
  static bool integer(const uint8_t **pp, const uint8_t *end, int prefix, uint64_t *out) {
    const uint8_t *p = *pp;
    if (p >= end) return false;
    const uint64_t mask = (1u << prefix) - 1;
    uint64_t v = *p++ & mask;
    if (v == mask) {
      int shift = 0;
      for (;;) {
        if (p >= end || shift > 56) return false;
        const uint8_t b = *p++;
        v += (uint64_t)(b & 0x7F) << shift;
        shift += 7;
        if (!(b & 0x80)) break;
      }
    }
    *pp = p;
    *out = v;
    return true;
  }
  static bool string(const uint8_t **pp, const uint8_t *end, std::string *out) {
    if (*pp >= end) return false;
    const bool huff = (**pp & 0x80) != 0;
    uint64_t len;
    if (!integer(pp, end, 7, &len) || (uint64_t)(end - *pp) < len) return false;
    if (huff) {
      if (!huffman().decode(*pp, (size_t)len, out)) return false;
    } else {
      out->assign(reinterpret_cast<const char *>(*pp), (size_t)len);
    }
    *pp += len;
    return true;
  }
  bool literal(const uint8_t **pp, const uint8_t *end, int prefix, std::string *n, std::string *v) {
    uint64_t idx;
    if (!integer(pp, end, prefix, &idx)) return false;
    if (idx == 0) {
      if (!string(pp, end, n)) return false;
    } else {
      std::string ignored;
      if (!lookup(idx, n, &ignored)) return false;
    }
    return string(pp, end, v);
  }
  bool lookup(uint64_t idx, std::string *n, std::string *v) const {
    if (idx == 0) return false;
    if (idx <= 61) {
      *n = kStaticTable[idx][0];
      *v = kStaticTable[idx][1];
      return true;
    }
    idx -= 62;
    if (idx >= dyn_.size()) return false;
    *n = dyn_[idx].first;
    *v = dyn_[idx].second;
    return true;
  }
  void insert(const std::string &n, const std::string &v) {
    size_ += n.size() + v.size() + 32;
    dyn_.emplace_front(n, v);
    evict();
  }
  void evict() {
    while (size_ > max_size_ && !dyn_.empty()) {
      size_ -= dyn_.back().first.size() + dyn_.back().second.size() + 32;
      dyn_.pop_back();
    }
  }
  std::deque<std::pair<std::string, std::string>> dyn_;
  size_t size_ = 0, max_size_ = 4096, settings_max_ = 4096;
};

inline void hpack_put_int(std::string *out, uint8_t first, int prefix, uint64_t v) {
  const uint64_t mask = (1u << prefix) - 1;
  if (v < mask) {
    out->push_back((char)(first | v));
    return;
  }
  out->push_back((char)(first | mask));
  v -= mask;
  while (v >= 128) {
    out->push_back((char)((v & 0x7F) | 0x80));
    v >>= 7;
  }
  out->push_back((char)v);
}
inline void hpack_put_str(std::string *out, const std::string &s) {
  hpack_put_int(out, 0x00, 7, s.size());  // H = 0: raw octets
  out->append(s);
}
// literal header field without indexing, new name (always valid, needs no shared state)
inline void hpack_put_header(std::string *out, const std::string &name, const std::string &value) {
  out->push_back(0x00);
  hpack_put_str(out, name);
  hpack_put_str(out, value);
}
inline void hpack_put_indexed(std::string *out, int idx) { hpack_put_int(out, 0x80, 7, (uint64_t)idx); }
// literal without indexing, name taken from the static table
inline void hpack_put_static_name(std::string *out, int name_idx, const std::string &value) {
  hpack_put_int(out, 0x00, 4, (uint64_t)name_idx);
  hpack_put_str(out, value);
}

// ------------------------------------------------------------------------------------ frames

enum : uint8_t { DATA = 0, HEADERS = 1, PRIORITY = 2, RST_STREAM = 3, SETTINGS = 4, PUSH_PROMISE = 5, PING = 6,
                 GOAWAY = 7, WINDOW_UPDATE = 8, CONTINUATION = 9 };
enum : uint8_t { F_END_STREAM = 0x1, F_ACK = 0x1, F_END_HEADERS = 0x4, F_PADDED = 0x8, F_PRIORITY = 0x20 };
static const char kPreface[] = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";

inline bool read_full(int fd, void *buf, size_t n) {
  uint8_t *p = static_cast<uint8_t *>(buf);
  while (n) {
    const ssize_t r = ::read(fd, p, n);
    if (r == 0) return false;
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += r;
    n -= (size_t)r;
  }
  return true;
}
inline bool write_full(int fd, const void *buf, size_t n) {
  const uint8_t *p = static_cast<const uint8_t *>(buf);
  while (n) {
    const ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += r;
    n -= (size_t)r;
  }
  return true;
}

struct Frame {
  uint8_t type = 0, flags = 0;
  uint32_t stream = 0;
  std::string payload;
};

inline bool read_frame(int fd, Frame *f, uint32_t max_len = 1u << 24) {
  uint8_t h[9];
  if (!read_full(fd, h, 9)) return false;
  const uint32_t len = ((uint32_t)h[0] << 16) | ((uint32_t)h[1] << 8) | h[2];
  if (len > max_len) return false;
  f->type = h[3];
  f->flags = h[4];
  f->stream = (((uint32_t)h[5] << 24) | ((uint32_t)h[6] << 16) | ((uint32_t)h[7] << 8) | h[8]) & 0x7FFFFFFFu;
  f->payload.resize(len);
  return len == 0 || read_full(fd, &f->payload[0], len);
}

inline std::string frame_bytes(uint8_t type, uint8_t flags, uint32_t stream, const void *payload, size_t len) {
  std::string out;
  out.resize(9 + len);
  out[0] = (char)(len >> 16);
  out[1] = (char)(len >> 8);
  out[2] = (char)len;
  out[3] = (char)type;
  out[4] = (char)flags;
  out[5] = (char)((stream >> 24) & 0x7F);
  out[6] = (char)(stream >> 16);
  out[7] = (char)(stream >> 8);
  out[8] = (char)stream;
  if (len) memcpy(&out[9], payload, len);
  return out;
}

inline std::string settings_payload(const std::vector<std::pair<uint16_t, uint32_t>> &kv) {
  std::string p;
  for (auto &e : kv) {
    p.push_back((char)(e.first >> 8));
    p.push_back((char)e.first);
    p.push_back((char)(e.second >> 24));
    p.push_back((char)(e.second >> 16));
    p.push_back((char)(e.second >> 8));
    p.push_back((char)e.second);
  }
  return p;
}

inline uint32_t be32(const uint8_t *p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

// strips padding / priority fields of HEADERS and DATA payloads; false on malformed padding
inline bool strip(const Frame &f, const uint8_t **p, size_t *n) {
  const uint8_t *b = reinterpret_cast<const uint8_t *>(f.payload.data());
  size_t len = f.payload.size(), pad = 0;
  if (f.flags & F_PADDED) {
    if (len < 1) return false;
    pad = b[0];
    b++;
    len--;
  }
  if (f.type == HEADERS && (f.flags & F_PRIORITY)) {
    if (len < 5) return false;
    b += 5;
    len -= 5;
  }
  if (pad > len) return false;
  *p = b;
  *n = len - pad;
  return true;
}

// ------------------------------------------------------------------------------------ gRPC server

class Connection;

// One server-side stream = one RPC.
struct Stream {
  uint32_t id = 0;
  std::string path;
  std::string body;      // request DATA (gRPC length-prefixed message)
  std::string hdr_block; // HEADERS + CONTINUATION fragments
  bool hdr_end_stream = false;
  bool too_large = false;  // request message beyond kMaxRecvMessage: answered RESOURCE_EXHAUSTED, bytes not kept
  std::string status_message;  // grpc-message of a non-OK status
  int64_t send_window = 65535;
  std::atomic<bool> cancelled{false};
  bool headers_sent = false;
};

// What a handler sees: the request message and a way to send response messages.
class Call {
 public:
  Call(std::shared_ptr<Connection> c, std::shared_ptr<Stream> s) : conn_(std::move(c)), stream_(std::move(s)) {}
  bool send_message(const void *msg, size_t len);  // false once the peer is gone / cancelled
  bool cancelled() const;
  const std::string &path() const { return stream_->path; }
  void set_status_message(const std::string &m) { stream_->status_message = m; }  // sent as grpc-message with a non-OK status

 private:
  friend class Connection;
  std::shared_ptr<Connection> conn_;
  std::shared_ptr<Stream> stream_;
};

// handler(call, request message) -> grpc-status (0 = OK). Runs on its own thread.
typedef std::function<int(Call &, const std::string &)> Handler;

// state that outlives the Server object for as long as any connection thread still runs
// Elastic worker pool for the RPC handlers: a call never waits behind another one (a streaming handler such as
// ListAndWatch occupies its worker for the life of the stream), but unary calls reuse parked threads instead of
// paying a thread creation + teardown each. A worker with nothing to do for 30 s exits.
class WorkerPool : public std::enable_shared_from_this<WorkerPool> {
 public:
  void submit(std::function<void()> fn) {
    std::unique_lock<std::mutex> lk(mu_);
    q_.push_back(std::move(fn));
    if (q_.size() > (size_t)idle_) {
      auto self = shared_from_this();
      lk.unlock();
      std::thread([self] { self->run(); }).detach();
    } else {
      cv_.notify_one();
    }
  }

 private:
  void run() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      if (q_.empty()) {
        idle_++;
        const bool got = cv_.wait_for(lk, std::chrono::seconds(30), [this] { return !q_.empty(); });
        idle_--;
        if (!got) return;
      }
      std::function<void()> fn = std::move(q_.front());
      q_.pop_front();
      lk.unlock();
      fn();
      fn = nullptr;  // drop the captures (connection, stream) before parking
      lk.lock();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  int idle_ = 0;
};

struct Shared {
  std::map<std::string, Handler> handlers;
  std::atomic<bool> stopping{false};
  std::shared_ptr<WorkerPool> pool = std::make_shared<WorkerPool>();
  // handlers currently executing: Server::stop() returns only once they have all returned, so whatever the
  // handlers capture (the plugin object) can be destroyed after it
  std::mutex active_mu;
  std::condition_variable active_cv;
  int active = 0;
};

class Connection : public std::enable_shared_from_this<Connection> {
 public:
  Connection(int fd, std::shared_ptr<Shared> sh) : fd_(fd), sh_(std::move(sh)) {}
  ~Connection() { ::close(fd_); }

  void serve() {  // reader loop; returns when the peer closes or violates the protocol
    char pre[24];
    if (!read_full(fd_, pre, 24) || memcmp(pre, kPreface, 24) != 0) return;
    {
      const std::string s = settings_payload({{0x3, 1024}, {0x4, kRecvWindow}, {0x5, 1u << 20}});
      std::string out = frame_bytes(SETTINGS, 0, 0, s.data(), s.size());
      // open the connection-level receive window as well (it starts at 65535 regardless of SETTINGS)
      uint8_t wu[4] = {0, 0, 0, 0};
      const uint32_t inc = kRecvWindow - 65535;
      wu[0] = (uint8_t)(inc >> 24), wu[1] = (uint8_t)(inc >> 16), wu[2] = (uint8_t)(inc >> 8), wu[3] = (uint8_t)inc;
      out += frame_bytes(WINDOW_UPDATE, 0, 0, wu, 4);
      if (!raw_write(out)) return;
    }
    Frame f;
    uint32_t continuation_of = 0;
    while (read_frame(fd_, &f)) {
      if (continuation_of && (f.type != CONTINUATION || f.stream != continuation_of)) break;
      switch (f.type) {
        case SETTINGS:
          if (f.flags & F_ACK) break;
          if (!on_settings(f)) goto done;
          break;
        case PING:
          if (!(f.flags & F_ACK) && f.payload.size() == 8)
            if (!raw_write(frame_bytes(PING, F_ACK, 0, f.payload.data(), 8))) goto done;
          break;
        case WINDOW_UPDATE:
          if (f.payload.size() == 4) on_window_update(f.stream, be32((const uint8_t *)f.payload.data()) & 0x7FFFFFFFu);
          break;
        case HEADERS:
        case CONTINUATION: {
          std::shared_ptr<Stream> s = f.type == HEADERS ? open_stream(f.stream) : find(f.stream);
          if (!s) goto done;
          if (f.type == HEADERS && open_streams() > kMaxConcurrentStreams) {  // beyond what SETTINGS advertised
            const uint8_t refused[4] = {0, 0, 0, 7};                           // REFUSED_STREAM: safe to retry
            erase(f.stream);
            if (!raw_write(frame_bytes(RST_STREAM, 0, f.stream, refused, 4))) goto done;
            if (!(f.flags & F_END_HEADERS)) goto done;  // its CONTINUATIONs would desynchronise HPACK: give up
            {
              const uint8_t *hp;
              size_t hn;
              Headers ignored;  // the block still has to pass through the decoder: HPACK state is per connection
              if (!strip(f, &hp, &hn) || !hpack_.decode(hp, hn, &ignored)) goto done;
            }
            break;
          }
          const uint8_t *p;
          size_t n;
          if (f.type == HEADERS) {
            if (!strip(f, &p, &n)) goto done;
            s->hdr_end_stream = (f.flags & F_END_STREAM) != 0;
          } else {
            p = (const uint8_t *)f.payload.data();
            n = f.payload.size();
          }
          s->hdr_block.append((const char *)p, n);
          if (s->hdr_block.size() > kMaxHeaderBlock) goto done;  // CONTINUATION flood: drop the connection
          if (f.flags & F_END_HEADERS) {
            continuation_of = 0;
            Headers hs;
            if (!hpack_.decode((const uint8_t *)s->hdr_block.data(), s->hdr_block.size(), &hs)) goto done;
            s->hdr_block.clear();
            for (auto &h : hs)
              if (h.first == ":path") s->path = h.second;
            if (s->hdr_end_stream) dispatch(s);
          } else {
            continuation_of = f.stream;
          }
          break;
        }
        case DATA: {
          std::shared_ptr<Stream> s = find(f.stream);
          const uint8_t *p;
          size_t n;
          if (!strip(f, &p, &n)) goto done;
          if (!f.payload.empty()) {  // give the credit back at once: requests on this surface are tiny
            uint8_t wu[4] = {(uint8_t)(f.payload.size() >> 24), (uint8_t)(f.payload.size() >> 16),
                             (uint8_t)(f.payload.size() >> 8), (uint8_t)f.payload.size()};
            std::string out = frame_bytes(WINDOW_UPDATE, 0, 0, wu, 4);
            if (s && !(f.flags & F_END_STREAM)) out += frame_bytes(WINDOW_UPDATE, 0, f.stream, wu, 4);
            if (!raw_write(out)) goto done;
          }
          if (!s) break;  // stream already finished/reset: ignore
          if (s->body.size() + n > kMaxRecvMessage + 5) {  // grpc-go's default MaxRecvMsgSize (4 MiB): refuse, do not buffer
            s->too_large = true;
            s->body.clear();
          } else if (!s->too_large) {
            s->body.append((const char *)p, n);
          }
          if (f.flags & F_END_STREAM) dispatch(s);
          break;
        }
        case RST_STREAM: {
          std::shared_ptr<Stream> s = find(f.stream);
          if (s) {
            s->cancelled = true;
            erase(f.stream);
            wake();
          }
          break;
        }
        case GOAWAY:
          goto done;
        default:
          break;  // PRIORITY, PUSH_PROMISE (never from a client), unknown: ignore
      }
    }
  done:
    closed_ = true;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto &e : streams_) e.second->cancelled = true;
    }
    wake();
    ::shutdown(fd_, SHUT_RDWR);
  }

  void go_away() {
    uint8_t p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    raw_write(frame_bytes(GOAWAY, 0, 0, p, 8));
    ::shutdown(fd_, SHUT_RDWR);
  }

  bool closed() const { return closed_.load(); }

 private:
  friend class Call;
  static constexpr uint32_t kRecvWindow = 4u << 20;
  static constexpr size_t kMaxRecvMessage = 4u << 20;    // grpc.NewServer() default (vendor/google.golang.org/grpc/server.go:53)
  static constexpr size_t kMaxHeaderBlock = 1u << 20;
  static constexpr size_t kMaxConcurrentStreams = 1024;  // what serve() advertises in SETTINGS
  size_t open_streams() {
    std::lock_guard<std::mutex> lk(mu_);
    return streams_.size();
  }

  bool raw_write(const std::string &bytes) {
    std::lock_guard<std::mutex> lk(wmu_);
    return !closed_ && write_full(fd_, bytes.data(), bytes.size());
  }
  void wake() {
    std::lock_guard<std::mutex> lk(mu_);
    cv_.notify_all();
  }
  std::shared_ptr<Stream> find(uint32_t id) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = streams_.find(id);
    return it == streams_.end() ? nullptr : it->second;
  }
  void erase(uint32_t id) {
    std::lock_guard<std::mutex> lk(mu_);
    streams_.erase(id);
  }
  std::shared_ptr<Stream> open_stream(uint32_t id) {
    if (id == 0 || (id & 1) == 0) return nullptr;
    std::lock_guard<std::mutex> lk(mu_);
    auto it = streams_.find(id);
    if (it != streams_.end()) return it->second;  // trailers of a client-streaming call: not used here
    if (id <= last_stream_) return nullptr;        // stream ids only grow (RFC 9113 5.1.1): connection error
    last_stream_ = id;
    auto s = std::make_shared<Stream>();
    s->id = id;
    s->send_window = peer_initial_window_;
    streams_[id] = s;
    return s;
  }
  bool on_settings(const Frame &f) {
    if (f.payload.size() % 6) return false;
    const uint8_t *p = (const uint8_t *)f.payload.data();
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (size_t i = 0; i + 6 <= f.payload.size(); i += 6) {
        const uint16_t id = (uint16_t)((p[i] << 8) | p[i + 1]);
        const uint32_t v = be32(p + i + 2);
        if (id == 0x4) {  // INITIAL_WINDOW_SIZE: shifts every open stream's window
          const int64_t delta = (int64_t)v - peer_initial_window_;
          peer_initial_window_ = v;
          for (auto &e : streams_) e.second->send_window += delta;
        } else if (id == 0x5) {
          peer_max_frame_ = v;
        }
      }
      cv_.notify_all();
    }
    return raw_write(frame_bytes(SETTINGS, F_ACK, 0, nullptr, 0));
  }
  void on_window_update(uint32_t stream, uint32_t inc) {
    std::lock_guard<std::mutex> lk(mu_);
    if (stream == 0) {
      conn_send_window_ += inc;
    } else {
      auto it = streams_.find(stream);
      if (it != streams_.end()) it->second->send_window += inc;
    }
    cv_.notify_all();
  }

  void dispatch(std::shared_ptr<Stream> s) {
    auto self = shared_from_this();
    sh_->pool->submit([self, s] {
      int status = 12;  // UNIMPLEMENTED
      std::string msg;
      bool ok_msg = false;
      if (s->too_large) {
        self->finish(s, 8);  // RESOURCE_EXHAUSTED, as grpc-go answers a message over its receive limit
        return;
      }
      if (s->body.size() >= 5 && s->body[0] == 0) {
        const uint32_t n = be32((const uint8_t *)s->body.data() + 1);
        if (s->body.size() == 5 + (size_t)n) {
          msg.assign(s->body, 5, n);
          ok_msg = true;
        }
      } else if (s->body.empty()) {
        ok_msg = true;  // no message at all: treat as an empty one
      }
      auto it = self->sh_->handlers.find(s->path);
      if (!ok_msg) {
        status = 13;  // INTERNAL: compressed or truncated message
      } else if (it != self->sh_->handlers.end()) {
        {
          std::lock_guard<std::mutex> lk(self->sh_->active_mu);
          self->sh_->active++;
        }
        if (!self->sh_->stopping) {
          Call call(self, s);
          status = it->second(call, msg);
        } else {
          status = 14;  // UNAVAILABLE: the server is going away
        }
        {
          std::lock_guard<std::mutex> lk(self->sh_->active_mu);
          self->sh_->active--;
          self->sh_->active_cv.notify_all();
        }
      }
      self->finish(s, status);
    });
  }

  bool send_headers_locked(Stream *s) {  // response HEADERS (wmu_ not held; sequence guarded by caller being the only writer of s)
    std::string block;
    hpack_put_indexed(&block, 8);                               // :status 200
    hpack_put_static_name(&block, 31, "application/grpc");      // content-type
    s->headers_sent = true;
    return raw_write(frame_bytes(HEADERS, F_END_HEADERS, s->id, block.data(), block.size()));
  }

  bool send_message(Stream *s, const void *msg, size_t len) {
    if (s->cancelled || closed_) return false;
    if (!s->headers_sent && !send_headers_locked(s)) return false;
    std::string body;
    body.resize(5 + len);
    body[0] = 0;
    body[1] = (char)(len >> 24);
    body[2] = (char)(len >> 16);
    body[3] = (char)(len >> 8);
    body[4] = (char)len;
    if (len) memcpy(&body[5], msg, len);
    size_t off = 0;
    while (off < body.size()) {
      size_t chunk;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] {
          return s->cancelled || closed_ || sh_->stopping.load() || (conn_send_window_ > 0 && s->send_window > 0);
        });
        if (s->cancelled || closed_ || sh_->stopping.load()) return false;
        chunk = body.size() - off;
        chunk = std::min<size_t>(chunk, peer_max_frame_);
        chunk = std::min<size_t>(chunk, (size_t)conn_send_window_);
        chunk = std::min<size_t>(chunk, (size_t)s->send_window);
        conn_send_window_ -= (int64_t)chunk;
        s->send_window -= (int64_t)chunk;
      }
      if (!raw_write(frame_bytes(DATA, 0, s->id, body.data() + off, chunk))) return false;
      off += chunk;
    }
    return true;
  }

  void finish(std::shared_ptr<Stream> s, int status) {
    if (!s->cancelled && !closed_) {
      std::string block;
      if (!s->headers_sent) {  // trailers-only response
        hpack_put_indexed(&block, 8);
        hpack_put_static_name(&block, 31, "application/grpc");
      }
      hpack_put_header(&block, "grpc-status", std::to_string(status));
      if (status != 0 && !s->status_message.empty()) {  // percent-encoded as the gRPC HTTP/2 spec asks
        std::string enc;
        for (unsigned char ch : s->status_message) {
          if (ch >= 0x20 && ch <= 0x7E && ch != '%') {
            enc.push_back((char)ch);
          } else {
            char e[4];
            snprintf(e, sizeof e, "%%%02X", ch);
            enc += e;
          }
        }
        hpack_put_header(&block, "grpc-message", enc);
      }
      raw_write(frame_bytes(HEADERS, F_END_HEADERS | F_END_STREAM, s->id, block.data(), block.size()));
    }
    erase(s->id);
  }

  int fd_;
  std::shared_ptr<Shared> sh_;
  std::atomic<bool> closed_{false};
  std::mutex wmu_;  // serialises writes to the socket
  std::mutex mu_;   // streams_, windows
  std::condition_variable cv_;
  std::map<uint32_t, std::shared_ptr<Stream>> streams_;
  HpackDecoder hpack_;
  int64_t conn_send_window_ = 65535;
  int64_t peer_initial_window_ = 65535;
  uint32_t peer_max_frame_ = 16384;
  uint32_t last_stream_ = 0;
};

inline bool Call::send_message(const void *msg, size_t len) { return conn_->send_message(stream_.get(), msg, len); }
inline bool Call::cancelled() const { return stream_->cancelled.load() || conn_->closed(); }

class Server {
 public:
  void handle(const std::string &path, Handler h) { sh_->handlers[path] = std::move(h); }  // before start()

  // net.Listen("unix", path) + go server.Serve (server.go:112-120)
  bool start(const std::string &socket_path, std::string *err) {
    lfd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (lfd_ < 0) return fail(err, "socket");
    sockaddr_un a;
    memset(&a, 0, sizeof a);
    a.sun_family = AF_UNIX;
    if (socket_path.size() >= sizeof a.sun_path) return fail(err, "socket path too long");
    strcpy(a.sun_path, socket_path.c_str());
    if (::bind(lfd_, (sockaddr *)&a, sizeof a) < 0) return fail(err, "bind");
    if (::listen(lfd_, 1024) < 0) return fail(err, "listen");
    sh_->stopping = false;
    acceptor_ = std::thread([this, lfd = lfd_] {  // the listener fd by value: stop() owns the member
      for (;;) {
        const int fd = ::accept4(lfd, nullptr, nullptr, SOCK_CLOEXEC);
        if (fd < 0) {
          if (errno == EINTR) continue;
          return;  // listener closed
        }
        auto c = std::make_shared<Connection>(fd, sh_);
        {
          std::lock_guard<std::mutex> lk(mu_);
          conns_.push_back(c);
          conns_.erase(std::remove_if(conns_.begin(), conns_.end(),
                                      [](const std::weak_ptr<Connection> &w) { return w.expired(); }),
                       conns_.end());
        }
        std::thread([c] { c->serve(); }).detach();
      }
    });
    return true;
  }

  // grpc.Server.Stop(): close the listener and every connection (server.go:141-143)
  void stop() {
    sh_->stopping = true;
    if (lfd_ >= 0) ::shutdown(lfd_, SHUT_RDWR);  // wakes accept4 with an error; the fd stays ours until it has returned
    if (acceptor_.joinable()) acceptor_.join();
    if (lfd_ >= 0) {
      ::close(lfd_);
      lfd_ = -1;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto &w : conns_)
        if (auto c = w.lock()) c->go_away();
      conns_.clear();
    }
    // handlers see their call cancelled (connection gone) and `stopping`; each is bounded by its own timeouts
    std::unique_lock<std::mutex> lk(sh_->active_mu);
    sh_->active_cv.wait(lk, [this] { return sh_->active == 0; });
  }

  ~Server() { stop(); }

 private:
  bool fail(std::string *err, const char *what) {
    if (err) *err = std::string(what) + ": " + strerror(errno);
    if (lfd_ >= 0) ::close(lfd_);
    lfd_ = -1;
    return false;
  }
  int lfd_ = -1;
  std::thread acceptor_;
  std::shared_ptr<Shared> sh_ = std::make_shared<Shared>();
  std::mutex mu_;
  std::vector<std::weak_ptr<Connection>> conns_;
};

// ------------------------------------------------------------------------------------ gRPC client (one unary call)

// Dial a unix socket, send one unary request, return the grpc-status (or -1 on transport failure).
// This is all Register needs (server.go:150-169: dial 5 s, one Register call, close).
inline int unary_call(const std::string &socket_path, const std::string &path, const std::string &request,
                      std::string *response, int timeout_s, std::string *err) {
  const int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) {
    if (err) *err = "socket";
    return -1;
  }
  struct Closer {
    int fd;
    ~Closer() { ::close(fd); }
  } closer{fd};
  timeval tv{timeout_s, 0};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
  sockaddr_un a;
  memset(&a, 0, sizeof a);
  a.sun_family = AF_UNIX;
  if (socket_path.size() >= sizeof a.sun_path) return -1;
  strcpy(a.sun_path, socket_path.c_str());
  // grpc.Dial(..., WithBlock(), WithTimeout(5s)) (server.go:90-104) keeps trying until the deadline: a kubelet that
  // has bound its socket but is not listening yet, or is just being restarted, is not an error
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(timeout_s);
  while (::connect(fd, (sockaddr *)&a, sizeof a) < 0) {
    const int e = errno;
    if ((e != ECONNREFUSED && e != ENOENT && e != EAGAIN && e != EINTR) || std::chrono::steady_clock::now() >= deadline) {
      if (err) *err = std::string("dial ") + socket_path + ": " + strerror(e);
      return -1;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
  }
  std::string out(kPreface, 24);
  out += frame_bytes(SETTINGS, 0, 0, nullptr, 0);
  std::string block;
  hpack_put_indexed(&block, 3);  // :method POST
  hpack_put_indexed(&block, 6);  // :scheme http
  hpack_put_static_name(&block, 4, path);
  hpack_put_static_name(&block, 1, "localhost");
  hpack_put_static_name(&block, 31, "application/grpc");
  hpack_put_header(&block, "te", "trailers");
  hpack_put_static_name(&block, 58, "gsbd-grpc/1");
  out += frame_bytes(HEADERS, F_END_HEADERS, 1, block.data(), block.size());
  std::string body;
  body.resize(5);
  body[0] = 0;
  body[1] = (char)(request.size() >> 24);
  body[2] = (char)(request.size() >> 16);
  body[3] = (char)(request.size() >> 8);
  body[4] = (char)request.size();
  body += request;
  out += frame_bytes(DATA, F_END_STREAM, 1, body.data(), body.size());
  if (!write_full(fd, out.data(), out.size())) {
    if (err) *err = "write failed";
    return -1;
  }
  HpackDecoder dec;
  Frame f;
  std::string hdr_block, data;
  int status = -1;
  bool end = false;
  while (!end && read_frame(fd, &f)) {
    if (f.type == SETTINGS && !(f.flags & F_ACK)) {
      const std::string ack = frame_bytes(SETTINGS, F_ACK, 0, nullptr, 0);
      write_full(fd, ack.data(), ack.size());
    } else if (f.type == PING && !(f.flags & F_ACK)) {
      const std::string ack = frame_bytes(PING, F_ACK, 0, f.payload.data(), f.payload.size());
      write_full(fd, ack.data(), ack.size());
    } else if ((f.type == HEADERS || f.type == CONTINUATION) && f.stream == 1) {
      const uint8_t *p;
      size_t n;
      if (f.type == HEADERS) {
        if (!strip(f, &p, &n)) break;
        if (f.flags & F_END_STREAM) end = true;
      } else {
        p = (const uint8_t *)f.payload.data();
        n = f.payload.size();
      }
      hdr_block.append((const char *)p, n);
      if (f.flags & F_END_HEADERS) {
        Headers hs;
        if (!dec.decode((const uint8_t *)hdr_block.data(), hdr_block.size(), &hs)) break;
        hdr_block.clear();
        for (auto &h : hs) {
          if (h.first == "grpc-status") status = atoi(h.second.c_str());
          if (h.first == "grpc-message" && err) *err = h.second;
        }
      } else {
        end = false;
      }
    } else if (f.type == DATA && f.stream == 1) {
      const uint8_t *p;
      size_t n;
      if (!strip(f, &p, &n)) break;
      data.append((const char *)p, n);
      if (f.flags & F_END_STREAM) end = true;
    } else if (f.type == RST_STREAM && f.stream == 1) {
      if (err) *err = "stream reset by peer";
      return -1;
    } else if (f.type == GOAWAY) {
      break;
    }
  }
  if (response && data.size() >= 5) response->assign(data, 5, std::string::npos);
  if (status < 0 && err && err->empty()) *err = "no grpc-status from peer";
  const std::string bye = frame_bytes(GOAWAY, 0, 0, "\0\0\0\0\0\0\0\0", 8);
  write_full(fd, bye.data(), bye.size());
  return status;
}

}  // namespace h2
