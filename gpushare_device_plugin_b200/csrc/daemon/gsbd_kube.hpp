// gsbd: kube API client configuration (podmanager.go kubeInit: kubeconfig or in-cluster) and calls
// Private to gsbd.cc (one translation unit): everything lives in an unnamed namespace.
#ifndef GSBD_KUBE_HPP_
#define GSBD_KUBE_HPP_
#include <stdlib.h>

#include <sstream>
#include <string>

#include "gsbd_flags.hpp"
#include "http_client.hpp"
#include "json.hpp"

namespace {

struct Kube {
  http::Client api;
  std::string node_name;

  // kubeInit (podmanager.go:29-57): $KUBECONFIG if the file exists, else in-cluster; NODE_NAME required.
  bool init(const Flags &f, std::string *err) {
    const char *nn = getenv("NODE_NAME");
    node_name = nn ? nn : "";
    if (node_name.empty()) {
      *err = "Please set env NODE_NAME";
      return false;
    }
    if (!f.kube_api_url.empty()) return api.configure(f.kube_api_url, "", "", true, 30, err);
    const char *kc = getenv("KUBECONFIG");
    if (kc && file_exists(kc)) {
      // first cluster / first user of the kubeconfig: server, token, CA (file or *-data), client cert/key
      // (file or *-data), insecure-skip-tls-verify — the forms kubeadm / cloud kubeconfigs use
      std::string server, token, ca;
      http::Client::TlsExtra extra;
      bool insecure = false;
      std::istringstream in(read_file(kc));
      std::string line;
      auto val = [](const std::string &l) {
        size_t c = l.find(':');
        std::string v = l.substr(c + 1);
        while (!v.empty() && (v.front() == ' ' || v.front() == '"' || v.front() == '\'')) v.erase(0, 1);
        while (!v.empty() && (v.back() == ' ' || v.back() == '"' || v.back() == '\'' || v.back() == '\r')) v.pop_back();
        return v;
      };
      auto b64 = [](const std::string &in) {
        std::string out;
        int acc = 0, bits = -8;
        for (unsigned char c : in) {
          int d = c >= 'A' && c <= 'Z' ? c - 'A' : c >= 'a' && c <= 'z' ? c - 'a' + 26 : c >= '0' && c <= '9' ? c - '0' + 52
                  : c == '+' ? 62 : c == '/' ? 63 : -1;
          if (d < 0) continue;
          acc = (acc << 6) | d;
          bits += 6;
          if (bits >= 0) {
            out.push_back((char)((acc >> bits) & 0xFF));
            bits -= 8;
          }
        }
        return out;
      };
      auto starts = [](const std::string &t, const char *k) { return t.compare(0, strlen(k), k) == 0; };
      while (std::getline(in, line)) {
        std::string t = line;
        t.erase(0, t.find_first_not_of(" -"));
        if (starts(t, "server:") && server.empty()) server = val(t);
        else if (starts(t, "token:") && token.empty()) token = val(t);
        else if (starts(t, "certificate-authority-data:") && extra.ca_pem.empty()) extra.ca_pem = b64(val(t));
        else if (starts(t, "certificate-authority:") && ca.empty()) ca = val(t);
        else if (starts(t, "client-certificate-data:") && extra.cert_pem.empty()) extra.cert_pem = b64(val(t));
        else if (starts(t, "client-key-data:") && extra.key_pem.empty()) extra.key_pem = b64(val(t));
        else if (starts(t, "client-certificate:") && extra.cert_file.empty()) extra.cert_file = val(t);
        else if (starts(t, "client-key:") && extra.key_file.empty()) extra.key_file = val(t);
        else if (starts(t, "insecure-skip-tls-verify:")) insecure = val(t) == "true";
      }
      if (server.empty()) {
        *err = std::string("no cluster server in ") + kc;
        return false;
      }
      return api.configure(server, token, ca, insecure, 30, err, &extra);
    }
    const char *h = getenv("KUBERNETES_SERVICE_HOST"), *p = getenv("KUBERNETES_SERVICE_PORT");
    if (!h || !p) {
      *err = "unable to load in-cluster configuration, KUBERNETES_SERVICE_HOST and KUBERNETES_SERVICE_PORT must be defined";
      return false;
    }
    const char *sa_dir = getenv("GSBD_SERVICEACCOUNT_DIR");  // tests; the pod's mount otherwise
    const std::string sa = sa_dir ? std::string(sa_dir) + "/" : "/var/run/secrets/kubernetes.io/serviceaccount/";
    token_file = sa + "token";
    std::string tok = read_file(token_file);
    while (!tok.empty() && (tok.back() == '\n' || tok.back() == '\r')) tok.pop_back();
    return api.configure(std::string("https://") + h + ":" + p, tok, sa + "ca.crt", false, 30, err);
  }
  std::string token_file;  // in-cluster only: the kubelet rewrites it before the token it holds expires

  // returns false + *err (= Status.message when the apiserver answered) on failure
  bool call(const std::string &method, const std::string &path, const std::string &body, const std::string &ctype,
            json::Value *out, std::string *err) {
    http::Response r;
    if (!api.request(method, path, body, ctype, &r, err)) return false;
    if (r.status == 401 && !token_file.empty()) {  // rotated service-account token: pick up the new one, once
      std::string tok = read_file(token_file);
      while (!tok.empty() && (tok.back() == '\n' || tok.back() == '\r')) tok.pop_back();
      if (!tok.empty() && tok != api.token()) {
        api.set_token(tok);
        r = http::Response();
        if (!api.request(method, path, body, ctype, &r, err)) return false;
      }
    }
    json::Value v;
    const bool parsed = json::parse(r.body, &v);
    if (r.status >= 400) {
      const json::Value *m = parsed ? v.get("message") : nullptr;
      *err = m ? m->str() : "HTTP " + std::to_string(r.status);
      return false;
    }
    if (!parsed) {
      *err = "undecodable response from apiserver";
      return false;
    }
    if (out) *out = std::move(v);
    return true;
  }
};

// resource.Quantity.Value(): integers with optional SI / binary suffix, fractions round up
uint64_t quantity_value(const json::Value &q) {
  std::string s = q.str();
  while (!s.empty() && s.back() == ' ') s.pop_back();
  static const struct { const char *suf; long double mult; } kSuf[] = {
      {"Ki", 1024.0L}, {"Mi", 1048576.0L}, {"Gi", 1073741824.0L}, {"Ti", 1099511627776.0L},
      {"Pi", 1125899906842624.0L}, {"Ei", 1152921504606846976.0L}, {"k", 1e3L}, {"M", 1e6L}, {"G", 1e9L},
      {"T", 1e12L}, {"P", 1e15L}, {"E", 1e18L}, {"m", 1e-3L}};
  long double mult = 1.0L;
  for (auto &e : kSuf) {
    const size_t n = strlen(e.suf);
    if (s.size() > n && s.compare(s.size() - n, n, e.suf) == 0) {
      mult = e.mult;
      s.resize(s.size() - n);
      break;
    }
  }
  const long double v = strtold(s.c_str(), nullptr) * mult;
  return v <= 0 ? 0 : (uint64_t)ceill(v - 1e-9L);
}


}  // namespace

#endif  // GSBD_KUBE_HPP_
