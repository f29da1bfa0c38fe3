// http_client.hpp — HTTP/1.1 keep-alive client for the kube-apiserver calls of the path (node GET,
// node-status PATCH, pod LIST, pod PATCH): plain TCP or TLS (OpenSSL), bearer token, Content-Length and
// chunked bodies, TCP_NODELAY like Go's net/http. Control-plane I/O: restated, not accelerated.
#pragma once

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <openssl/err.h>
#include <openssl/pem.h>
#include <openssl/x509v3.h>
#include <openssl/ssl.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace http {

struct Response {
  int status = 0;
  std::string body;
};

class Conn {
 public:
  ~Conn() { close(); }
  bool open(const std::string &host, int port, SSL_CTX *tls, const std::string &sni, int timeout_s, std::string *err,
            bool verify_name = true) {
    addrinfo hints, *res = nullptr;
    memset(&hints, 0, sizeof hints);
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || !res) {
      *err = "dial tcp: lookup " + host + " failed";
      return false;
    }
    for (addrinfo *a = res; a; a = a->ai_next) {
      fd_ = ::socket(a->ai_family, a->ai_socktype | SOCK_CLOEXEC, a->ai_protocol);
      if (fd_ < 0) continue;
      timeval tv{timeout_s, 0};
      setsockopt(fd_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
      setsockopt(fd_, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
      if (::connect(fd_, a->ai_addr, a->ai_addrlen) == 0) break;
      ::close(fd_);
      fd_ = -1;
    }
    freeaddrinfo(res);
    if (fd_ < 0) {
      *err = "dial tcp " + host + ":" + std::to_string(port) + ": " + strerror(errno);
      return false;
    }
    int one = 1;
    setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    if (tls) {
      ssl_ = SSL_new(tls);
      SSL_set_fd(ssl_, fd_);
      // the certificate must be FOR the endpoint we dialled: DNS name or (in-cluster: 10.x service IP) IP SAN
      unsigned char ipbuf[16];
      const bool is_ip = inet_pton(AF_INET, sni.c_str(), ipbuf) == 1 || inet_pton(AF_INET6, sni.c_str(), ipbuf) == 1;
      if (!sni.empty() && !is_ip) SSL_set_tlsext_host_name(ssl_, sni.c_str());
      if (verify_name && !sni.empty()) {
        X509_VERIFY_PARAM *vp = SSL_get0_param(ssl_);
        if (is_ip) X509_VERIFY_PARAM_set1_ip_asc(vp, sni.c_str());
        else X509_VERIFY_PARAM_set1_host(vp, sni.c_str(), 0);
      }
      if (SSL_connect(ssl_) != 1) {
        const long vr = SSL_get_verify_result(ssl_);
        *err = "tls handshake with " + host + " failed" +
               (vr != X509_V_OK ? std::string(": x509: ") + X509_verify_cert_error_string(vr) : std::string());
        close();
        return false;
      }
    }
    return true;
  }
  void close() {
    if (ssl_) {
      SSL_free(ssl_);
      ssl_ = nullptr;
    }
    if (fd_ >= 0) ::close(fd_);
    fd_ = -1;
    buf_.clear();
  }
  bool write_all(const std::string &s) {
    size_t off = 0;
    while (off < s.size()) {
      const int n = ssl_ ? SSL_write(ssl_, s.data() + off, (int)(s.size() - off))
                         : (int)::send(fd_, s.data() + off, s.size() - off, MSG_NOSIGNAL);
      if (n <= 0) return false;
      off += (size_t)n;
    }
    return true;
  }
  bool fill() {
    char tmp[16384];
    const int n = ssl_ ? SSL_read(ssl_, tmp, sizeof tmp) : (int)::recv(fd_, tmp, sizeof tmp, 0);
    if (n <= 0) return false;
    buf_.append(tmp, (size_t)n);
    return true;
  }
  bool read_line(std::string *line) {
    for (;;) {
      const size_t p = buf_.find("\r\n");
      if (p != std::string::npos) {
        line->assign(buf_, 0, p);
        buf_.erase(0, p + 2);
        return true;
      }
      if (!fill()) return false;
    }
  }
  bool read_n(size_t n, std::string *out) {
    while (buf_.size() < n)
      if (!fill()) return false;
    out->append(buf_, 0, n);
    buf_.erase(0, n);
    return true;
  }
  // Streaming variant for watches: parses the status line + headers, then hands every '\n'-terminated line of
  // the (chunked or close-delimited) body to `on_line` as it arrives. Returns when the server ends the stream,
  // the connection drops, on_line returns false, or abort() is called from another thread.
  template <typename F, typename H>
  bool read_stream(int *status, H on_headers, F on_line) {
    std::string line;
    if (!read_line(&line) || line.size() < 12) return false;
    *status = atoi(line.c_str() + 9);
    bool chunked = false;
    while (read_line(&line)) {
      if (line.empty()) break;
      std::string lower = line;
      for (auto &c : lower) c = (char)tolower((unsigned char)c);
      if (lower.compare(0, 18, "transfer-encoding:") == 0 && lower.find("chunked") != std::string::npos) chunked = true;
    }
    on_headers(*status);
    std::string pending;
    auto feed = [&](const std::string &data) -> bool {
      pending += data;
      size_t nl;
      while ((nl = pending.find('\n')) != std::string::npos) {
        std::string one = pending.substr(0, nl);
        pending.erase(0, nl + 1);
        if (!one.empty() && !on_line(one)) return false;
      }
      return true;
    };
    if (*status >= 400) {  // error body: small, read what is there
      std::string body;
      while (fill()) {
      }
      body.swap(buf_);
      on_line(body);
      return false;
    }
    if (chunked) {
      for (;;) {
        if (!read_line(&line)) return false;
        const size_t n = (size_t)strtoul(line.c_str(), nullptr, 16);
        if (n == 0) return true;
        std::string chunk;
        if (!read_n(n, &chunk) || !read_line(&line)) return false;
        if (!feed(chunk)) return true;
      }
    }
    for (;;) {
      if (!buf_.empty()) {
        std::string chunk;
        chunk.swap(buf_);
        if (!feed(chunk)) return true;
      }
      if (!fill()) return true;
    }
  }
  void abort() {  // from another thread: wakes a blocked read
    if (fd_ >= 0) ::shutdown(fd_, SHUT_RDWR);
  }
  void set_read_timeout(int seconds) {
    timeval tv{seconds, 0};
    if (fd_ >= 0) setsockopt(fd_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  }

  bool read_response(Response *r, bool *keep_alive) {
    std::string line;
    if (!read_line(&line) || line.size() < 12) return false;
    r->status = atoi(line.c_str() + 9);
    long content_length = -1;
    bool chunked = false;
    *keep_alive = true;
    while (read_line(&line)) {
      if (line.empty()) break;
      std::string lower = line;
      for (auto &c : lower) c = (char)tolower((unsigned char)c);
      if (lower.compare(0, 15, "content-length:") == 0) content_length = atol(line.c_str() + 15);
      if (lower.compare(0, 18, "transfer-encoding:") == 0 && lower.find("chunked") != std::string::npos) chunked = true;
      if (lower.compare(0, 11, "connection:") == 0 && lower.find("close") != std::string::npos) *keep_alive = false;
    }
    r->body.clear();
    if (chunked) {
      for (;;) {
        if (!read_line(&line)) return false;
        const size_t n = (size_t)strtoul(line.c_str(), nullptr, 16);
        if (n == 0) {
          while (read_line(&line) && !line.empty()) {
          }
          return true;
        }
        if (!read_n(n, &r->body) || !read_line(&line)) return false;
      }
    }
    if (content_length >= 0) return read_n((size_t)content_length, &r->body);
    *keep_alive = false;  // body delimited by close
    while (fill()) {
    }
    r->body.swap(buf_);
    return true;
  }

 private:
  int fd_ = -1;
  SSL *ssl_ = nullptr;
  std::string buf_;
};

class Client {
 public:
  Client() = default;
  Client(const Client &) = delete;
  Client &operator=(const Client &) = delete;
  ~Client() {
    idle_.clear();
    if (ctx_) SSL_CTX_free(ctx_);
  }
  // base_url: http://host:port or https://host:port
  // optional TLS material given as PEM text (kubeconfig *-data fields) or files (client cert for mTLS)
  struct TlsExtra {
    std::string ca_pem, cert_pem, key_pem, cert_file, key_file;
  };
  bool configure(const std::string &base_url, const std::string &token, const std::string &ca_file, bool insecure,
                 int timeout_s, std::string *err, const TlsExtra *extra = nullptr) {
    token_ = token;
    timeout_s_ = timeout_s;
    insecure_ = insecure;
    {
      std::lock_guard<std::mutex> lk(mu_);
      idle_.clear();  // connections made under the previous configuration
    }
    if (ctx_) {
      SSL_CTX_free(ctx_);
      ctx_ = nullptr;
    }
    tls_ = false;
    std::string rest;
    if (base_url.compare(0, 8, "https://") == 0) {
      tls_ = true;
      rest = base_url.substr(8);
    } else if (base_url.compare(0, 7, "http://") == 0) {
      rest = base_url.substr(7);
    } else {
      *err = "unsupported URL " + base_url;
      return false;
    }
    const size_t slash = rest.find('/');
    if (slash != std::string::npos) rest.resize(slash);
    const size_t colon = rest.rfind(':');
    host_ = colon == std::string::npos ? rest : rest.substr(0, colon);
    port_ = colon == std::string::npos ? (tls_ ? 443 : 80) : atoi(rest.c_str() + colon + 1);
    if (tls_) {
      ctx_ = SSL_CTX_new(TLS_client_method());
      if (!ctx_) {
        *err = "SSL_CTX_new failed";
        return false;
      }
      if (insecure) {
        SSL_CTX_set_verify(ctx_, SSL_VERIFY_NONE, nullptr);
      } else {
        SSL_CTX_set_verify(ctx_, SSL_VERIFY_PEER, nullptr);
        if (extra && !extra->ca_pem.empty()) {
          BIO *bio = BIO_new_mem_buf(extra->ca_pem.data(), (int)extra->ca_pem.size());
          X509 *x = nullptr;
          int n = 0;
          while ((x = PEM_read_bio_X509(bio, nullptr, nullptr, nullptr)) != nullptr) {
            X509_STORE_add_cert(SSL_CTX_get_cert_store(ctx_), x);
            X509_free(x);
            n++;
          }
          BIO_free(bio);
          ERR_clear_error();
          if (n == 0) {
            *err = "certificate-authority-data holds no PEM certificate";
            return false;
          }
        } else if (!ca_file.empty()) {
          if (SSL_CTX_load_verify_locations(ctx_, ca_file.c_str(), nullptr) != 1) {
            *err = "cannot load CA " + ca_file;
            return false;
          }
        } else {
          SSL_CTX_set_default_verify_paths(ctx_);
        }
      }
      if (extra && (!extra->cert_pem.empty() || !extra->cert_file.empty())) {  // client certificate (mTLS)
        bool ok = true;
        if (!extra->cert_pem.empty()) {
          BIO *cb = BIO_new_mem_buf(extra->cert_pem.data(), (int)extra->cert_pem.size());
          X509 *x = PEM_read_bio_X509(cb, nullptr, nullptr, nullptr);
          BIO_free(cb);
          BIO *kb = BIO_new_mem_buf(extra->key_pem.data(), (int)extra->key_pem.size());
          EVP_PKEY *k = PEM_read_bio_PrivateKey(kb, nullptr, nullptr, nullptr);
          BIO_free(kb);
          ok = x && k && SSL_CTX_use_certificate(ctx_, x) == 1 && SSL_CTX_use_PrivateKey(ctx_, k) == 1;
          if (x) X509_free(x);
          if (k) EVP_PKEY_free(k);
        } else {
          ok = SSL_CTX_use_certificate_chain_file(ctx_, extra->cert_file.c_str()) == 1 &&
               SSL_CTX_use_PrivateKey_file(ctx_, extra->key_file.c_str(), SSL_FILETYPE_PEM) == 1;
        }
        if (!ok) {
          *err = "cannot load the client certificate / key";
          return false;
        }
      }
    }
    return true;
  }

  // bearer token, replaceable while requests are in flight (a projected service-account token is rotated on disk)
  std::string token() {
    std::lock_guard<std::mutex> lk(mu_);
    return token_;
  }
  void set_token(const std::string &t) {
    std::lock_guard<std::mutex> lk(mu_);
    token_ = t;
  }

  // A dedicated connection for a long-lived GET (watch): returns the connection with the request already sent.
  std::unique_ptr<Conn> open_stream(const std::string &path, int read_timeout_s, std::string *err) {
    std::unique_ptr<Conn> c(new Conn());
    if (!c->open(host_, port_, ctx_, host_, timeout_s_, err, !insecure_)) return nullptr;
    c->set_read_timeout(read_timeout_s);
    std::string req = "GET " + path + " HTTP/1.1\r\nHost: " + host_ + "\r\nAccept: application/json\r\n";
    const std::string tok = token();
    if (!tok.empty()) req += "Authorization: Bearer " + tok + "\r\n";
    req += "\r\n";
    if (!c->write_all(req)) {
      *err = "GET " + path + ": connection failed";
      return nullptr;
    }
    return c;
  }

  // One request; a stale pooled connection is retried once on a fresh one.
  bool request(const std::string &method, const std::string &path, const std::string &body,
               const std::string &content_type, Response *out, std::string *err) {
    std::string req = method + " " + path + " HTTP/1.1\r\nHost: " + host_ + "\r\nAccept: application/json\r\n";
    const std::string tok = token();
    if (!tok.empty()) req += "Authorization: Bearer " + tok + "\r\n";
    if (!content_type.empty()) req += "Content-Type: " + content_type + "\r\n";
    if (!body.empty() || method == "PATCH" || method == "POST" || method == "PUT")
      req += "Content-Length: " + std::to_string(body.size()) + "\r\n";
    req += "\r\n";
    req += body;  // headers and body in one write
    for (int attempt = 0; attempt < 2; attempt++) {
      std::unique_ptr<Conn> c = take();
      const bool pooled = c != nullptr;
      if (!c) {
        c.reset(new Conn());
        if (!c->open(host_, port_, ctx_, host_, timeout_s_, err, !insecure_)) return false;
      }
      bool keep = false;
      if (c->write_all(req) && c->read_response(out, &keep)) {
        if (keep) give(std::move(c));
        return true;
      }
      if (!pooled) {
        *err = method + " " + path + ": connection failed";
        return false;
      }
    }
    *err = method + " " + path + ": connection failed";
    return false;
  }

 private:
  std::unique_ptr<Conn> take() {
    std::lock_guard<std::mutex> lk(mu_);
    if (idle_.empty()) return nullptr;
    std::unique_ptr<Conn> c = std::move(idle_.back());
    idle_.pop_back();
    return c;
  }
  void give(std::unique_ptr<Conn> c) {
    std::lock_guard<std::mutex> lk(mu_);
    if (idle_.size() < 64) idle_.push_back(std::move(c));
  }
  std::string host_, token_;
  int port_ = 0, timeout_s_ = 30;
  bool tls_ = false, insecure_ = false;
  SSL_CTX *ctx_ = nullptr;
  std::mutex mu_;
  std::vector<std::unique_ptr<Conn>> idle_;
};

}  // namespace http
