// gsbd: constants of the wire contract (const.go, v1beta1/constants.go) and glog-shaped logging
// Private to gsbd.cc (one translation unit): everything lives in an unnamed namespace.
#ifndef GSBD_LOG_HPP_
#define GSBD_LOG_HPP_
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>

#include "../../../include/gpushare_b200.h"

namespace {

// ---------------------------------------------------------------- constants (const.go, v1beta1/constants.go)
const char kResourceName[] = "aliyun.com/gpu-mem";
const char kResourceCount[] = "aliyun.com/gpu-count";
const char kDevicePluginPath[] = "/var/lib/kubelet/device-plugins/";
const char kServerSockName[] = "aliyungpushare.sock";
const char kOptimisticLockErrorMsg[] =
    "the object has been modified; please apply your changes to the latest version and try again";
const char kEnvResourceIndex[] = "ALIYUN_COM_GPU_MEM_IDX";
const char kEnvAssignedFlag[] = "ALIYUN_COM_GPU_MEM_ASSIGNED";
const char kEnvResourceAssumeTime[] = "ALIYUN_COM_GPU_MEM_ASSUME_TIME";
const char kEnvNodeLabelForDisableCGPU[] = "cgpu.disable.isolation";


// ---------------------------------------------------------------- logging (glog-shaped, stderr)
// The reference logs synchronously from inside Allocate's critical section (>= 6 glog lines per call at --v=5,
// SURVEY §8 a12). Here a line is formatted by the caller and handed to one writer thread; the RPC path never
// waits for stderr (a container runtime's log pipe). Order is preserved; warnings and errors, and everything
// when GSBD_SYNC_LOG=1, are written before logf returns; log_flush() runs before every exit.
int g_v = 0;
class AsyncLog {
 public:
  void write(const char *line, size_t n, bool sync) {
    std::unique_lock<std::mutex> lk(mu_);
    buf_.append(line, n);
    if (sync || !running_ || buf_.size() > (1u << 20)) {  // also the back-pressure path: never grow without bound
      drain_locked();
      return;
    }
    if (idle_) cv_.notify_one();
  }
  void start() {
    std::lock_guard<std::mutex> lk(mu_);
    if (running_ || getenv("GSBD_SYNC_LOG")) return;
    running_ = true;
    th_ = std::thread([this] { run(); });
  }
  void flush() {  // stop the writer and write what is left; logging stays usable (synchronous) afterwards
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (!running_) {
        drain_locked();
        return;
      }
      running_ = false;
      cv_.notify_one();
    }
    th_.join();
    std::lock_guard<std::mutex> lk(mu_);
    drain_locked();
  }

 private:
  void drain_locked() {
    if (buf_.empty()) return;
    fwrite(buf_.data(), 1, buf_.size(), stderr);
    fflush(stderr);
    buf_.clear();
  }
  void run() {
    std::string out;
    std::unique_lock<std::mutex> lk(mu_);
    while (running_) {
      if (buf_.empty()) {
        idle_ = true;
        cv_.wait(lk, [this] { return !running_ || !buf_.empty(); });
        idle_ = false;
      }
      out.swap(buf_);
      lk.unlock();
      if (!out.empty()) {
        fwrite(out.data(), 1, out.size(), stderr);
        fflush(stderr);
        out.clear();
      }
      lk.lock();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::string buf_;
  bool running_ = false, idle_ = false;
  std::thread th_;
} g_log;
void log_flush() { g_log.flush(); }

void logf(char sev, const char *fmt, ...) {
  char line[2200];
  timeval tv;
  gettimeofday(&tv, nullptr);
  tm t;
  localtime_r(&tv.tv_sec, &t);
  int n = snprintf(line, 64, "%c%02d%02d %02d:%02d:%02d.%06ld %7d gsbd] ", sev, t.tm_mon + 1, t.tm_mday, t.tm_hour,
                   t.tm_min, t.tm_sec, (long)tv.tv_usec, (int)getpid());
  va_list ap;
  va_start(ap, fmt);
  const int m = vsnprintf(line + n, sizeof line - (size_t)n - 1, fmt, ap);
  va_end(ap);
  n += m < 0 ? 0 : std::min(m, (int)sizeof line - n - 2);
  line[n++] = '\n';
  g_log.write(line, (size_t)n, sev != 'I');
}
#define INFO(...) logf('I', __VA_ARGS__)
#define WARN(...) logf('W', __VA_ARGS__)
#define VLOG(n, ...)                 \
  do {                               \
    if (g_v >= (n)) logf('I', __VA_ARGS__); \
  } while (0)

std::string last_error() {
  char buf[512];
  gsb_last_error(buf, sizeof buf);
  return buf;
}


}  // namespace

#endif  // GSBD_LOG_HPP_
