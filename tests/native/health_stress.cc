// Host-side concurrency of the C ABI on a box without a driver: the health event queue (inject / wait / stop from
// many threads) and the lifecycle entry points racing each other. Built with -fsanitize=thread against a
// TSan-instrumented build of the library by tools/sanitize.sh; also a plain smoke test (exit 0, counts printed).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "gpushare_b200.h"

int main(int argc, char **argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 2.0;
  std::atomic<bool> stop{false}, stop_inject{false};
  std::atomic<long> injected{0}, received{0}, timeouts{0}, stopped{0}, lifecycle{0};
  std::vector<std::thread> ts;
  for (int i = 0; i < 4; i++)
    ts.emplace_back([&, i] {
      gsb_event ev;
      memset(&ev, 0, sizeof ev);
      snprintf(ev.uuid, sizeof ev.uuid, "GPU-%08d-0000-0000-0000-000000000000", i);
      ev.etype = 8;
      while (!stop_inject) {
        ev.edata = 31 + (uint64_t)(injected % 50);
        if (gsb_health_inject(&ev) == GSB_OK) injected++;
        if ((injected & 63) == 0) std::this_thread::sleep_for(std::chrono::microseconds(200));
      }
    });
  for (int i = 0; i < 4; i++)
    ts.emplace_back([&] {
      gsb_event ev;
      while (!stop) {
        const int rc = gsb_health_wait(5, &ev);
        if (rc == GSB_OK) received++;
        else if (rc == GSB_ERR_TIMEOUT) timeouts++;
        else if (rc == GSB_ERR_STOPPED) stopped++;
      }
    });
  ts.emplace_back([&] {  // the plugin's Stop(): wakes every waiter, queue usable again afterwards
    while (!stop) {
      gsb_health_stop();
      std::this_thread::sleep_for(std::chrono::milliseconds(7));
    }
  });
  for (int i = 0; i < 2; i++)
    ts.emplace_back([&] {  // lifecycle and inventory entry points with no driver present: errors, never a crash
      char buf[256];
      uint32_t n = 0;
      gsb_device_info info;
      while (!stop) {
        gsb_init();
        gsb_device_count(&n);
        gsb_device_info_get(0, &info);
        gsb_last_error(buf, sizeof buf);
        gsb_health_start(100, 1 << 20);
        gsb_shutdown();
        lifecycle++;
      }
    });
  std::this_thread::sleep_for(std::chrono::milliseconds((int)(seconds * 1000)));
  stop_inject = true;  // producers first; the consumers then drain what is queued (every event exactly once)
  for (int i = 0; i < 4; i++) ts[(size_t)i].join();
  for (int spin = 0; spin < 400 && received.load() < injected.load(); spin++)
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
  stop = true;
  for (size_t i = 4; i < ts.size(); i++) ts[i].join();
  printf("injected %ld received %ld timeouts %ld stopped %ld lifecycle %ld\n", injected.load(), received.load(),
         timeouts.load(), stopped.load(), lifecycle.load());
  return received.load() > 0 && lifecycle.load() > 0 ? 0 : 1;
}
