// h2_selftest — tiny gRPC server/client built on h2.hpp, driven by tests/test_native_h2.py against
// grpcio (the C-core HTTP/2 stack): unary echo, large responses (flow control), server streaming,
// cancellation, and the one-shot client used for Registration/Register.
//   h2_selftest serve <socket>        serves until stdin closes
//   h2_selftest call <socket> <path> <hex request>   prints "<status> <hex response>"
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <string>

#include "h2.hpp"
#include "json.hpp"

static std::string unhex(const char *s) {
  std::string o;
  for (size_t i = 0; s[i] && s[i + 1]; i += 2) o.push_back((char)strtol(std::string(s + i, 2).c_str(), nullptr, 16));
  return o;
}

int main(int argc, char **argv) {
  if (argc >= 3 && std::string(argv[1]) == "serve") {
    h2::Server srv;
    srv.handle("/test.Echo/Unary", [](h2::Call &c, const std::string &req) {
      c.send_message(req.data(), req.size());
      return 0;
    });
    // request = decimal "<count> <bytes>": stream <count> messages of <bytes> bytes each
    srv.handle("/test.Echo/Stream", [](h2::Call &c, const std::string &req) {
      int count = 0, bytes = 0;
      sscanf(req.c_str(), "%d %d", &count, &bytes);
      for (int i = 0; i < count; i++) {
        std::string m((size_t)bytes, (char)('a' + i % 26));
        if (!c.send_message(m.data(), m.size())) return 1;  // cancelled
      }
      return 0;
    });
    srv.handle("/test.Echo/Forever", [](h2::Call &c, const std::string &) {
      while (!c.cancelled()) {
        if (!c.send_message("tick", 4)) break;
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
      }
      fprintf(stderr, "forever: cancelled\n");
      return 1;
    });
    srv.handle("/test.Echo/Fail", [](h2::Call &, const std::string &) { return 5; });
    std::string err;
    if (!srv.start(argv[2], &err)) {
      fprintf(stderr, "%s\n", err.c_str());
      return 1;
    }
    printf("ready\n");
    fflush(stdout);
    while (getchar() != EOF) {
    }
    srv.stop();
    return 0;
  }
  if (argc >= 5 && std::string(argv[1]) == "call") {
    std::string resp, err;
    const int st = h2::unary_call(argv[2], argv[3], unhex(argv[4]), &resp, 5, &err);
    printf("%d ", st);
    for (unsigned char ch : resp) printf("%02x", ch);
    printf(" %s\n", err.c_str());
    return st == 0 ? 0 : 2;
  }
  if (argc >= 3 && std::string(argv[1]) == "hpack") {  // decode header blocks in sequence with ONE decoder
    h2::HpackDecoder dec;
    for (int i = 2; i < argc; i++) {
      const std::string block = unhex(argv[i]);
      h2::Headers hs;
      if (!dec.decode((const uint8_t *)block.data(), block.size(), &hs)) {
        printf("ERROR\n");
        return 1;
      }
      for (auto &h : hs) printf("%s: %s\n", h.first.c_str(), h.second.c_str());
      printf("--\n");
    }
    return 0;
  }
  if (argc >= 2 && std::string(argv[1]) == "json") {  // stdin JSON -> canonical re-serialisation (or "INVALID")
    std::string in, out;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) in.append(buf, n);
    json::Value v;
    if (!json::parse(in, &v)) {
      printf("INVALID\n");
      return 1;
    }
    json::dump(v, &out);
    fwrite(out.data(), 1, out.size(), stdout);
    return 0;
  }
  fprintf(stderr, "usage: h2_selftest serve <socket> | call <socket> <path> <hex request> | hpack <hex>...\n");
  return 64;
}
