/*
 * gsb_wire.cc — host-only half of libgpushare_b200.so: slice arithmetic, fake device IDs, the
 * device-plugin v1beta1 wire encoders/decoder and the Allocate decision. No CUDA in this file.
 *
 * Mirrors (reference paths):
 *   pkg/gpu/nvidia/nvidia.go:26-45                      IDs and MiB->slice arithmetic
 *   vendor/k8s.io/kubernetes/pkg/kubelet/apis/deviceplugin/v1beta1/api.pb.go
 *        :730-767  RegisterRequest.MarshalTo     :794-812  ListAndWatchResponse.MarshalTo
 *        :824-843  Device.MarshalTo              :940-957  ContainerAllocateRequest.MarshalTo
 *        :969-987  AllocateResponse.MarshalTo    :999-1062 ContainerAllocateResponse.MarshalTo
 *   pkg/gpu/nvidia/allocate.go:24-198, podutils.go:27-35,78-119, podmanager.go:215-262
 */
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <string_view>
#include <unordered_set>
#include <vector>

#include "../../include/gpushare_b200.h"

namespace {

inline size_t varint_size(uint64_t v) {  // sovApi
  size_t n = 1;
  while (v >= 0x80) {
    v >>= 7;
    n++;
  }
  return n;
}

inline uint8_t *put_varint(uint8_t *p, uint64_t v) {  // encodeVarintApi
  while (v >= 0x80) {
    *p++ = (uint8_t)(v | 0x80);
    v >>= 7;
  }
  *p++ = (uint8_t)v;
  return p;
}

inline uint8_t *put_bytes(uint8_t *p, uint8_t tag, const char *s, size_t n) {
  *p++ = tag;
  p = put_varint(p, n);
  memcpy(p, s, n);
  return p + n;
}

inline size_t dec_len(uint32_t v) {
  return v < 10 ? 1 : v < 100 ? 2 : v < 1000 ? 3 : v < 10000 ? 4 : v < 100000 ? 5 : v < 1000000 ? 6
         : v < 10000000 ? 7 : v < 100000000 ? 8 : v < 1000000000 ? 9 : 10;
}

inline char *put_dec(char *p, uint64_t v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

std::string dec(uint64_t v) {
  char b[24];
  return std::string(b, put_dec(b, v));
}

const char kHealthy[] = "Healthy";      // pluginapi.Healthy   (v1beta1/constants.go)
const char kUnhealthy[] = "Unhealthy";  // pluginapi.Unhealthy

// ---- proto3 reader with the accept/reject behaviour of gogo's generated Unmarshal ----------------------------
// (api.pb.go:2141-2300 for AllocateRequest / ContainerAllocateRequest, skipApi at :2991-3089): what grpc-go would
// refuse with "error unmarshalling request" is refused here too, what it skips is skipped.

struct Reader {
  const uint8_t *p, *end;
  bool varint(uint64_t *v) {  // <= 10 bytes, else ErrIntOverflowApi; running out = io.ErrUnexpectedEOF
    uint64_t r = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) return false;
      const uint8_t b = *p++;
      r |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) {
        *v = r;
        return true;
      }
    }
    return false;
  }
  // skipApi: p stands on the KEY of an unknown field; consumes the whole field. Groups recurse (depth-bounded here:
  // Go would recurse without limit, nothing legitimate nests 64 groups).
  bool skip_field(int depth = 0) {
    if (depth > 64) return false;
    uint64_t key, n;
    if (!varint(&key)) return false;
    switch (key & 7) {
      case 0: return varint(&n);
      case 1: if (end - p < 8) return false; p += 8; return true;
      case 2:
        if (!varint(&n) || (int64_t)n < 0 || (uint64_t)(end - p) < n) return false;  // int(length) < 0: ErrInvalidLength
        p += n;
        return true;
      case 3:
        for (;;) {
          const uint8_t *start = p;
          uint64_t inner;
          if (!varint(&inner)) return false;
          if ((inner & 7) == 4) return true;
          p = start;
          if (!skip_field(depth + 1)) return false;
        }
      case 4: return true;
      case 5: if (end - p < 4) return false; p += 4; return true;
      default: return false;  // "proto: illegal wireType"
    }
  }
};

// One message whose only known field is number 1, length-delimited (both request messages have that shape).
template <typename F>
bool decode_field1_message(const uint8_t *b, size_t len, F on_field1) {
  Reader r{b, b + len};
  while (r.p < r.end) {
    const uint8_t *pre = r.p;
    uint64_t key;
    if (!r.varint(&key)) return false;
    const int32_t field = (int32_t)(uint32_t)(key >> 3);  // int32(wire >> 3), wrap included
    const uint32_t wire = (uint32_t)(key & 7);
    if (wire == 4 || field <= 0) return false;  // "wiretype end group for non-group" / "illegal tag"
    if (field == 1) {
      if (wire != 2) return false;  // "wrong wireType"
      uint64_t n;
      if (!r.varint(&n) || (int64_t)n < 0 || (uint64_t)(r.end - r.p) < n) return false;
      if (!on_field1(r.p, (size_t)n)) return false;
      r.p += n;
    } else {
      r.p = pre;
      if (!r.skip_field()) return false;
    }
  }
  return true;
}

// AllocateRequest -> number of devicesIDs per container request
bool decode_allocate_request(const uint8_t *req, size_t len, std::vector<uint64_t> *per_container) {
  return decode_field1_message(req, len, [&](const uint8_t *c, size_t n) {
    uint64_t ids = 0;
    if (!decode_field1_message(c, n, [&](const uint8_t *, size_t) {
          ids++;
          return true;
        }))
      return false;
    per_container->push_back(ids);
    return true;
  });
}

struct Env {
  const char *k;
  std::string v;
};

// ContainerAllocateResponse with only envs set; keys must arrive sorted
size_t container_response_size(const std::vector<Env> &envs) {
  size_t n = 0;
  for (const Env &e : envs) {
    const size_t kl = strlen(e.k), vl = e.v.size();
    const size_t map_size = 1 + kl + varint_size(kl) + 1 + vl + varint_size(vl);
    n += 1 + varint_size(map_size) + map_size;
  }
  return n;
}

uint8_t *put_container_response(uint8_t *p, const std::vector<Env> &envs) {
  for (const Env &e : envs) {
    const size_t kl = strlen(e.k), vl = e.v.size();
    const size_t map_size = 1 + kl + varint_size(kl) + 1 + vl + varint_size(vl);
    *p++ = 0x0a;
    p = put_varint(p, map_size);
    p = put_bytes(p, 0x0a, e.k, kl);
    p = put_bytes(p, 0x12, e.v.data(), vl);
  }
  return p;
}

int encode_allocate_response(const std::vector<std::vector<Env>> &containers, uint8_t *resp, size_t cap,
                             size_t *resp_len) {
  size_t total = 0;
  for (const auto &c : containers) {
    const size_t sz = container_response_size(c);
    total += 1 + varint_size(sz) + sz;
  }
  *resp_len = total;
  if (!resp || cap < total) return GSB_ERR_BUFFER_TOO_SMALL;
  uint8_t *p = resp;
  for (const auto &c : containers) {
    *p++ = 0x0a;
    p = put_varint(p, container_response_size(c));
    p = put_container_response(p, c);
  }
  return GSB_OK;
}

// env keys (const.go:24-32), listed in sorted order
const char kEnvContainer[] = "ALIYUN_COM_GPU_MEM_CONTAINER";
const char kEnvDev[] = "ALIYUN_COM_GPU_MEM_DEV";
const char kEnvIdx[] = "ALIYUN_COM_GPU_MEM_IDX";
const char kEnvPod[] = "ALIYUN_COM_GPU_MEM_POD";
const char kEnvCgpuDisable[] = "CGPU_DISABLE";
const char kEnvNvGpu[] = "NVIDIA_VISIBLE_DEVICES";

// buildErrResponse (allocate.go:24-39)
std::vector<std::vector<Env>> err_envs(const gsb_allocate_ctx *ctx, const std::vector<uint64_t> &per_container,
                                       uint64_t pod_req) {
  std::vector<std::vector<Env>> out;
  for (uint64_t n : per_container) {
    std::vector<Env> e;
    e.push_back({kEnvContainer, dec(n)});
    e.push_back({kEnvDev, dec(ctx->slices)});
    e.push_back({kEnvIdx, "-1"});
    e.push_back({kEnvPod, dec(pod_req)});
    e.push_back({kEnvNvGpu, "no-gpu-has-" + dec(pod_req) + (ctx->unit_gib ? "GiB" : "MiB") + "-to-run"});
    out.push_back(std::move(e));
  }
  return out;
}

// allocate.go:113-128 (matched pod) and :160-175 (single-GPU shortcut)
std::vector<std::vector<Env>> ok_envs(const gsb_allocate_ctx *ctx, const std::vector<uint64_t> &per_container,
                                      uint64_t pod_req, const std::string &visible, uint64_t idx) {
  std::vector<std::vector<Env>> out;
  for (uint64_t n : per_container) {
    std::vector<Env> e;
    e.push_back({kEnvContainer, dec(n)});
    e.push_back({kEnvDev, dec(ctx->slices)});
    e.push_back({kEnvIdx, dec(idx)});
    e.push_back({kEnvPod, dec(pod_req)});
    if (ctx->disable_cgpu_isolation) e.push_back({kEnvCgpuDisable, "true"});
    e.push_back({kEnvNvGpu, visible});
    out.push_back(std::move(e));
  }
  return out;
}

// isGPUMemoryAssumedPod (podutils.go:78-119)
inline bool is_assumed(const gsb_pod &p) {
  if (p.gpu_mem_limit == 0) return false;
  if (!p.has_assume_time) return false;
  return p.has_assigned && p.assigned_is_false;
}

}  // namespace

extern "C" {

uint32_t gsb_slices(uint64_t total_mib, int unit_gib) {
  // setGPUMemory: v := raw; if metric == GiBPrefix { v = raw / 1024 }   (nvidia.go:34-41)
  return (uint32_t)(unit_gib ? total_mib / 1024ull : total_mib);
}

int gsb_fake_device_id(const char *uuid, uint32_t j, char *buf, size_t cap) {
  if (!uuid || !buf) return GSB_ERR_INVALID_ARGUMENT;
  const size_t ul = strlen(uuid), need = ul + 3 + dec_len(j);
  if (cap < need + 1) return GSB_ERR_BUFFER_TOO_SMALL;
  memcpy(buf, uuid, ul);
  memcpy(buf + ul, "-_-", 3);
  char *e = put_dec(buf + ul + 3, j);
  *e = 0;
  return (int)need;
}

int gsb_real_device_id(const char *fake_id, char *buf, size_t cap) {
  if (!fake_id || !buf) return GSB_ERR_INVALID_ARGUMENT;
  const char *sep = strstr(fake_id, "-_-");
  const size_t n = sep ? (size_t)(sep - fake_id) : strlen(fake_id);
  if (cap < n + 1) return GSB_ERR_BUFFER_TOO_SMALL;
  memcpy(buf, fake_id, n);
  buf[n] = 0;
  return (int)n;
}

int64_t gsb_encode_list_and_watch(const char *const *uuids, uint32_t n_gpus, uint32_t slices,
                                  const uint8_t *unhealthy_bits, uint8_t *buf, size_t cap) {
  if (n_gpus && !uuids) return GSB_ERR_INVALID_ARGUMENT;
  // pass 1: size
  size_t total = 0;
  for (uint32_t g = 0; g < n_gpus; g++) {
    if (!uuids[g]) return GSB_ERR_INVALID_ARGUMENT;
    const size_t ul = strlen(uuids[g]);
    for (uint32_t j = 0; j < slices; j++) {
      const size_t id_len = ul + 3 + dec_len(j);
      const size_t bit = (size_t)g * slices + j;
      const bool bad = unhealthy_bits && (unhealthy_bits[bit >> 3] >> (bit & 7) & 1);
      const size_t hl = bad ? sizeof(kUnhealthy) - 1 : sizeof(kHealthy) - 1;
      const size_t dev = 1 + varint_size(id_len) + id_len + 1 + 1 + hl;
      total += 1 + varint_size(dev) + dev;
    }
  }
  if (!buf) return (int64_t)total;
  if (cap < total) return GSB_ERR_BUFFER_TOO_SMALL;
  // pass 2: bytes
  uint8_t *p = buf;
  for (uint32_t g = 0; g < n_gpus; g++) {
    const char *u = uuids[g];
    const size_t ul = strlen(u);
    for (uint32_t j = 0; j < slices; j++) {
      const size_t id_len = ul + 3 + dec_len(j);
      const size_t bit = (size_t)g * slices + j;
      const bool bad = unhealthy_bits && (unhealthy_bits[bit >> 3] >> (bit & 7) & 1);
      const char *h = bad ? kUnhealthy : kHealthy;
      const size_t hl = bad ? sizeof(kUnhealthy) - 1 : sizeof(kHealthy) - 1;
      const size_t dev = 1 + varint_size(id_len) + id_len + 1 + 1 + hl;
      *p++ = 0x0a;  // ListAndWatchResponse.devices
      p = put_varint(p, dev);
      *p++ = 0x0a;  // Device.ID
      p = put_varint(p, id_len);
      memcpy(p, u, ul);
      p += ul;
      *p++ = '-';
      *p++ = '_';
      *p++ = '-';
      p = (uint8_t *)put_dec((char *)p, j);
      p = put_bytes(p, 0x12, h, hl);  // Device.health
    }
  }
  return (int64_t)(p - buf);
}

int64_t gsb_encode_register_request(const char *version, const char *endpoint, const char *resource_name,
                                    uint8_t *buf, size_t cap) {
  const char *f[3] = {version ? version : "", endpoint ? endpoint : "", resource_name ? resource_name : ""};
  const uint8_t tags[3] = {0x0a, 0x12, 0x1a};
  size_t total = 0;
  for (int i = 0; i < 3; i++) {
    const size_t n = strlen(f[i]);
    if (n) total += 1 + varint_size(n) + n;  // proto3: empty strings are omitted
  }
  if (!buf) return (int64_t)total;
  if (cap < total) return GSB_ERR_BUFFER_TOO_SMALL;
  uint8_t *p = buf;
  for (int i = 0; i < 3; i++) {
    const size_t n = strlen(f[i]);
    if (n) p = put_bytes(p, tags[i], f[i], n);
  }
  return (int64_t)(p - buf);
}

}  // extern "C"

namespace {

// sort.Sort as Go 1.10 runs it (the reference builds with golang:1.10: Dockerfile:1, .travis.yml:3-4) over the
// candidate pods with the reference's NON-STRICT Less (podmanager.go:256-258: assume[i] <= assume[j]).
// Third-party dependency absent from the reference tree: Go standard library, package sort, go1.10, src/sort/sort.go
// (insertionSort, siftDown, heapSort, medianOfThree, doPivot, quickSort, Sort, maxDepth), restated from its published
// source; the same restatement lives in oracle/wire_oracle.py::go110_sort and the two are held against each other
// (tests/test_allocate.py, tests/golden/allocate_cases.json). `v` holds indices into `pods`; Less/Swap act on positions.
struct Go110Sort {
  const gsb_pod *pods;
  std::vector<uint32_t> &v;

  bool less(long i, long j) const { return pods[v[(size_t)i]].assume_time <= pods[v[(size_t)j]].assume_time; }
  void swap(long i, long j) { std::swap(v[(size_t)i], v[(size_t)j]); }

  void insertion_sort(long a, long b) {
    for (long i = a + 1; i < b; i++)
      for (long j = i; j > a && less(j, j - 1); j--) swap(j, j - 1);
  }
  void sift_down(long lo, long hi, long first) {
    long root = lo;
    for (;;) {
      long child = 2 * root + 1;
      if (child >= hi) return;
      if (child + 1 < hi && less(first + child, first + child + 1)) child++;
      if (!less(first + root, first + child)) return;
      swap(first + root, first + child);
      root = child;
    }
  }
  void heap_sort(long a, long b) {
    const long first = a, lo = 0, hi = b - a;
    for (long i = (hi - 1) / 2; i >= 0; i--) sift_down(i, hi, first);
    for (long i = hi - 1; i >= 0; i--) {
      swap(first, first + i);
      sift_down(lo, i, first);
    }
  }
  void median_of_three(long m1, long m0, long m2) {
    if (less(m1, m0)) swap(m1, m0);
    if (less(m2, m1)) {
      swap(m2, m1);
      if (less(m1, m0)) swap(m1, m0);
    }
  }
  void do_pivot(long lo, long hi, long *midlo, long *midhi) {
    const long m = (long)((unsigned long)(lo + hi) >> 1);
    if (hi - lo > 40) {  // Tukey's ninther
      const long s = (hi - lo) / 8;
      median_of_three(lo, lo + s, lo + 2 * s);
      median_of_three(m, m - s, m + s);
      median_of_three(hi - 1, hi - 1 - s, hi - 1 - 2 * s);
    }
    median_of_three(lo, m, hi - 1);
    const long pivot = lo;
    long a = lo + 1, c = hi - 1;
    for (; a < c && less(a, pivot); a++) {
    }
    long b = a;
    for (;;) {
      for (; b < c && !less(pivot, b); b++) {
      }
      for (; b < c && less(pivot, c - 1); c--) {
      }
      if (b >= c) break;
      swap(b, c - 1);
      b++;
      c--;
    }
    bool protect = hi - c < 5;
    if (!protect && hi - c < (hi - lo) / 4) {
      int dups = 0;
      if (!less(pivot, hi - 1)) {
        swap(c, hi - 1);
        c++;
        dups++;
      }
      if (!less(b - 1, pivot)) {
        b--;
        dups++;
      }
      if (!less(m, pivot)) {
        swap(m, b - 1);
        b--;
        dups++;
      }
      protect = dups > 1;
    }
    if (protect) {
      for (;;) {
        for (; a < b && !less(b - 1, pivot); b--) {
        }
        for (; a < b && less(a, pivot); a++) {
        }
        if (a >= b) break;
        swap(a, b - 1);
        a++;
        b--;
      }
    }
    swap(pivot, b - 1);
    *midlo = b - 1;
    *midhi = c;
  }
  void quick_sort(long a, long b, int max_depth) {
    while (b - a > 12) {
      if (max_depth == 0) {
        heap_sort(a, b);
        return;
      }
      max_depth--;
      long mlo, mhi;
      do_pivot(a, b, &mlo, &mhi);
      if (mlo - a < b - mhi) {
        quick_sort(a, mlo, max_depth);
        a = mhi;
      } else {
        quick_sort(mhi, b, max_depth);
        b = mlo;
      }
    }
    if (b - a > 1) {
      for (long i = a + 6; i < b; i++)
        if (less(i, i - 6)) swap(i, i - 6);
      insertion_sort(a, b);
    }
  }
  void sort() {
    const long n = (long)v.size();
    int depth = 0;
    for (long i = n; i > 0; i >>= 1) depth++;
    quick_sort(0, n, depth * 2);
  }
};

}  // namespace

extern "C" {

int gsb_allocate(const gsb_allocate_ctx *ctx, const gsb_pod *pods, uint32_t n_pods, const uint8_t *req,
                 size_t req_len, uint8_t *resp, size_t resp_cap, size_t *resp_len, int32_t *pod_index,
                 uint32_t *pod_req_gpu) {
  if (!ctx || (!pods && n_pods) || (!req && req_len) || !resp_len) return GSB_ERR_INVALID_ARGUMENT;
  if (pod_index) *pod_index = -1;
  std::vector<uint64_t> per_container;
  if (!decode_allocate_request(req, req_len, &per_container)) return GSB_ERR_MALFORMED;
  uint64_t pod_req = 0;  // allocate.go:54-56
  for (uint64_t n : per_container) pod_req += n;
  if (pod_req_gpu) *pod_req_gpu = (uint32_t)pod_req;

  // getPendingPodsInNode's dedupe by UID (podmanager.go:162-212: the first pod of a UID wins, whether or not it
  // is a candidate), then the candidate filter. The set is an open-addressing table of (hash, index) kept
  // across calls on this thread; a hash hit is confirmed on the bytes, so the result is exact.
  std::vector<uint32_t> cand;
  if (ctx->pods_unique) {
    cand.reserve(n_pods);
    for (uint32_t i = 0; i < n_pods; i++)
      if (pods[i].on_node && is_assumed(pods[i])) cand.push_back(i);
  } else {
    thread_local std::vector<uint64_t> slots;  // hash << 32 | (index + 1); 0 = empty
    size_t cap = 64;
    while (cap < (size_t)n_pods * 2) cap <<= 1;
    slots.assign(cap, 0);
    cand.reserve(n_pods);
    for (uint32_t i = 0; i < n_pods; i++) {
      if (!pods[i].on_node) continue;
      const char *u = pods[i].uid ? pods[i].uid : "";
      uint64_t h = 1469598103934665603ull;  // FNV-1a
      size_t len = 0;
      for (; u[len]; len++) h = (h ^ (unsigned char)u[len]) * 1099511628211ull;
      const uint32_t tag = (uint32_t)(h >> 32) | 1u;
      size_t at = (size_t)h & (cap - 1);
      bool dup = false;
      while (slots[at]) {
        if ((uint32_t)(slots[at] >> 32) == tag) {
          const char *o = pods[(uint32_t)slots[at] - 1].uid;
          if (strcmp(o ? o : "", u) == 0) {
            dup = true;
            break;
          }
        }
        at = (at + 1) & (cap - 1);
      }
      if (dup) continue;
      slots[at] = (uint64_t)tag << 32 | (uint64_t)(i + 1);
      if (is_assumed(pods[i])) cand.push_back(i);
    }
  }
  // makePodOrderdByAge: sort.Sort with Less = (t[i] <= t[j])  (podmanager.go:241-262), then the first pod in
  // that order whose request equals this one (allocate.go:78-88). With the non-strict Less, the order of pods with
  // EQUAL assume-times is whatever Go 1.10's sort.Sort leaves behind, so that algorithm is restated whole
  // (Go110Sort above). It is only RUN when it can matter: the output is sorted by time whatever the tie order, so
  // if exactly one matching candidate carries the smallest matching time it is the answer — one O(n) pass, no sort
  // (every call of config 4/5, where timestamps are distinct nanoseconds). Only a tie for that smallest time among
  // matching candidates needs the real permutation.
  int32_t found = -1;
  uint32_t tied = 0;
  for (uint32_t c : cand) {
    if (pods[c].gpu_mem_limit != pod_req) continue;
    if (found < 0 || pods[c].assume_time < pods[found].assume_time) {
      found = (int32_t)c;
      tied = 1;
    } else if (pods[c].assume_time == pods[found].assume_time) {
      tied++;
    }
  }
  if (tied > 1) {
    Go110Sort sorter{pods, cand};
    sorter.sort();
    found = -1;
    for (uint32_t c : cand)
      if (pods[c].gpu_mem_limit == pod_req) {
        found = (int32_t)c;
        break;
      }
  }

  int kind;
  std::vector<std::vector<Env>> envs;
  if (found >= 0) {
    int64_t id = pods[found].gpu_idx;  // getGPUIDFromPodAnnotation (podutils.go:37-61)
    const char *uuid = nullptr;
    if (id >= 0) {  // GetDeviceNameByIndex: minor -> UUID (server.go:72-83)
      for (uint32_t g = 0; g < ctx->n_gpus; g++)
        if (ctx->minors[g] == (uint64_t)id) uuid = ctx->uuids[g];
      if (!uuid) id = -1;
    }
    if (id < 0) {
      kind = GSB_ALLOC_ERR_RESPONSE;  // allocate.go:108-110
      envs = err_envs(ctx, per_container, pod_req);
    } else {
      kind = GSB_ALLOC_MATCHED;
      envs = ok_envs(ctx, per_container, pod_req, dec((uint64_t)id), (uint64_t)id);  // %v of int
      if (pod_index) *pod_index = found;
    }
  } else if (ctx->n_gpus == 1) {  // allocate.go:151-177
    kind = GSB_ALLOC_SINGLE_GPU;
    envs = ok_envs(ctx, per_container, pod_req, ctx->uuids[0], ctx->minors[0]);
  } else {
    kind = GSB_ALLOC_ERR_RESPONSE;  // allocate.go:179-184
    envs = err_envs(ctx, per_container, pod_req);
  }
  const int rc = encode_allocate_response(envs, resp, resp_cap, resp_len);
  return rc ? rc : kind;
}

int gsb_allocate_err_response(const gsb_allocate_ctx *ctx, const uint8_t *req, size_t req_len, uint8_t *resp,
                              size_t resp_cap, size_t *resp_len) {
  if (!ctx || (!req && req_len) || !resp_len) return GSB_ERR_INVALID_ARGUMENT;
  std::vector<uint64_t> per_container;
  if (!decode_allocate_request(req, req_len, &per_container)) return GSB_ERR_MALFORMED;
  uint64_t pod_req = 0;
  for (uint64_t n : per_container) pod_req += n;
  return encode_allocate_response(err_envs(ctx, per_container, pod_req), resp, resp_cap, resp_len);
}

int gsb_patch_assigned_body(uint64_t now_unix_ns, char *buf, size_t cap) {
  if (!buf) return GSB_ERR_INVALID_ARGUMENT;
  // json.Marshal of the nested map: keys sorted, no whitespace (podutils.go:27-35)
  const int n = snprintf(buf, cap,
                         "{\"metadata\":{\"annotations\":{\"ALIYUN_COM_GPU_MEM_ASSIGNED\":\"true\","
                         "\"ALIYUN_COM_GPU_MEM_ASSUME_TIME\":\"%llu\"}}}",
                         (unsigned long long)now_unix_ns);
  if (n < 0 || (size_t)n >= cap) return GSB_ERR_BUFFER_TOO_SMALL;
  return n;
}

}  // extern "C"
