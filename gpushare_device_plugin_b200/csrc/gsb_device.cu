/*
 * gsb_device.cu — device layer of libgpushare_b200.so: lifecycle, inventory, arena, probe, cycle,
 * health events. Host code; the kernels live in hbm_probe_sm100a.cu.
 *
 * Replaces the reference's go-nvml use (vendor/github.com/NVIDIA/gpu-monitoring-tools/bindings/go/
 * nvml/{nvml.go,bindings.go,nvml_dl.c}) for the inventory+health path:
 *   - nvml.Init (nvml_dl.c:21-28: dlopen + nvmlInit_v2)             -> gsb_init
 *   - nvml.NewDevice's 11 getters + sysfs read (nvml.go:297-359)     -> 4 NVML calls per device
 *     (handle, UUID, minor, MemoryInfo v1) + 2 CUDA driver calls that cross-check identity; the
 *     eight getters whose values the plugin discards (name, power, PCI, BAR1, link gen/width,
 *     clocks, NUMA) are not made
 *   - total MiB = nvmlMemory_t.total / (1024*1024) (bindings.go:346-349): SAME call, SAME struct, so
 *     the slice count is bit-exact by construction. cuDeviceTotalMem is 762 839 040 B smaller on a
 *     B200 (profiles/envprobe_r01.txt: 178 vs 179 GiB) and is reported only as information.
 *   - RegisterEventForDevice x fake devices (nvidia.go:104-117; O(S*N^2) NVML calls) -> one
 *     nvmlDeviceRegisterEvents per GPU
 *   - WaitForEvent (bindings.go:134-146) -> gsb_health_wait over a queue fed by the XID thread and
 *     by the active HBM prober
 *
 * libcuda.so.1 and libnvidia-ml.so.1 are dlopen'ed at gsb_init, never linked: the library loads on
 * a GPU-less builder (symbol check in tests/test_abi.py) and reports GSB_ERR_LIBRARY_NOT_FOUND there.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvml.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include "gsb_internal.h"

namespace {

// ------------------------------------------------------------------ error text

thread_local char tl_error[512];

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof tl_error, fmt, ap);
  va_end(ap);
}

uint64_t now_ns() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

// ------------------------------------------------------------------ dlopen'ed entry points

struct DriverApi {
  void *lib = nullptr;
  CUresult (*cuInit)(unsigned) = nullptr;
  CUresult (*cuDeviceGet)(CUdevice *, int) = nullptr;
  CUresult (*cuDeviceGetUuid)(CUuuid *, CUdevice) = nullptr;
  CUresult (*cuDeviceTotalMem)(size_t *, CUdevice) = nullptr;
  CUresult (*cuMemGetInfo)(size_t *, size_t *) = nullptr;
  CUresult (*cuMemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *,
                                            CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*cuMemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*cuMemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*cuMemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *,
                          unsigned long long) = nullptr;
  CUresult (*cuMemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*cuMemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*cuGetErrorString)(CUresult, const char **) = nullptr;
};

struct NvmlApi {
  void *lib = nullptr;
  nvmlReturn_t (*init)(void) = nullptr;
  nvmlReturn_t (*shutdown)(void) = nullptr;
  const char *(*errorString)(nvmlReturn_t) = nullptr;
  nvmlReturn_t (*getCount)(unsigned *) = nullptr;
  nvmlReturn_t (*handleByIndex)(unsigned, nvmlDevice_t *) = nullptr;
  nvmlReturn_t (*getUUID)(nvmlDevice_t, char *, unsigned) = nullptr;
  nvmlReturn_t (*getMinor)(nvmlDevice_t, unsigned *) = nullptr;
  nvmlReturn_t (*getMemoryInfo)(nvmlDevice_t, nvmlMemory_t *) = nullptr;  // v1 struct: the reference's
  nvmlReturn_t (*getPciInfo)(nvmlDevice_t, nvmlPciInfo_t *) = nullptr;
  nvmlReturn_t (*eventSetCreate)(nvmlEventSet_t *) = nullptr;
  nvmlReturn_t (*registerEvents)(nvmlDevice_t, unsigned long long, nvmlEventSet_t) = nullptr;
  nvmlReturn_t (*eventSetWait)(nvmlEventSet_t, nvmlEventData_t *, unsigned) = nullptr;
  nvmlReturn_t (*eventSetFree)(nvmlEventSet_t) = nullptr;
};

template <typename F>
bool load_sym(void *lib, F &fn, const char *name, const char *alt = nullptr) {
  fn = reinterpret_cast<F>(dlsym(lib, name));
  if (!fn && alt) fn = reinterpret_cast<F>(dlsym(lib, alt));
  return fn != nullptr;
}

// ------------------------------------------------------------------ state

constexpr uint64_t kGranuleBytes = 64ull << 20;  // generation-table granularity
constexpr uint32_t kGranuleShiftWords = 22;      // log2(64 MiB / 16 B)

struct Chunk {
  CUmemGenericAllocationHandle handle;
  size_t size;
};

struct Device {
  nvmlDevice_t nvml{};
  int ordinal = -1;
  CUdevice cudev = 0;
  char uuid[GSB_UUID_BUFFER_SIZE] = {0};
  char bus_id[GSB_BUSID_BUFFER_SIZE] = {0};
  uint32_t minor = 0;
  uint32_t sm_count = 0, cc_major = 0, cc_minor = 0;
  uint64_t cuda_total = 0;

  std::mutex mu;  // serialises arena + launches of this device
  bool ready = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  gsb_kernel_out *out_host = nullptr, *out_dev = nullptr;
  gsb_partial *partials = nullptr;
  unsigned int *ticket = nullptr;
  unsigned long long *tile_counter = nullptr;
  uint32_t *seed_table = nullptr;  // device copy of gen[]
  uint32_t launch_seq = 0;

  CUdeviceptr va = 0;
  size_t va_size = 0;
  std::vector<Chunk> chunks;
  uint64_t arena_bytes = 0;
  std::map<uint64_t, gsb_launch_geom> geoms;  // (op, variant, grid request) -> resolved launch geometry
  std::vector<uint32_t> gen;  // host mirror of seed_table: generation that last wrote each granule
  uint32_t next_gen = 1;
  bool faulted = false;  // prober: sticky

  // persistent worker (gsb_cycle_all / gsb_probe_all): one host thread per device
  std::thread worker;
  std::mutex wmu;
  std::condition_variable wcv;
  std::function<void()> job;
  bool job_ready = false, job_done = false, worker_quit = false;
  std::atomic<bool> job_flag{false}, done_flag{false}, quit_flag{false};  // lock-free mirrors for the spin phase
};

struct Global {
  // every device-touching entry point holds this shared; gsb_init/gsb_shutdown take it exclusively, so a
  // shutdown waits for calls still running on other threads (e.g. a start-up walk) instead of freeing
  // devices under them
  std::shared_mutex api_mu;
  std::mutex mu;
  bool inited = false;
  DriverApi cu;
  NvmlApi ml;
  std::vector<std::unique_ptr<Device>> devs;  // NVML index order

  // health
  std::mutex hmu;
  std::condition_variable hcv;
  std::deque<gsb_event> events;
  bool health_running = false;
  std::atomic<bool> health_stop{false};
  std::atomic<uint32_t> recovery_cycles{0};
  uint64_t stop_gen = 0;  // bumped by every gsb_health_stop: wakes all gsb_health_wait callers
  std::vector<std::thread> health_threads;
  nvmlEventSet_t event_set{};
  bool have_event_set = false;
};

Global G;

const char *cu_err(CUresult r) {
  const char *s = nullptr;
  if (G.cu.cuGetErrorString && G.cu.cuGetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown CUDA driver error";
}

#define CU_TRY(expr)                                                      \
  do {                                                                    \
    CUresult _r = (expr);                                                 \
    if (_r != CUDA_SUCCESS) {                                             \
      set_error("%s -> %d (%s)", #expr, (int)_r, cu_err(_r));             \
      return GSB_ERR_DRIVER;                                              \
    }                                                                     \
  } while (0)

#define RT_TRY(expr)                                                      \
  do {                                                                    \
    cudaError_t _r = (expr);                                              \
    if (_r != cudaSuccess) {                                              \
      set_error("%s -> %d (%s)", #expr, (int)_r, cudaGetErrorString(_r)); \
      return _r == cudaErrorMemoryAllocation ? GSB_ERR_OUT_OF_MEMORY : GSB_ERR_DRIVER; \
    }                                                                     \
  } while (0)

// "nvml: <nvmlErrorString>" is the reference's error text (bindings.go:52-58)
#define ML_TRY(expr)                                                      \
  do {                                                                    \
    nvmlReturn_t _r = (expr);                                             \
    if (_r != NVML_SUCCESS) {                                             \
      set_error("nvml: %s", G.ml.errorString ? G.ml.errorString(_r) : "?"); \
      return GSB_ERR_NVML;                                                \
    }                                                                     \
  } while (0)

void format_uuid(const CUuuid &u, char out[GSB_UUID_BUFFER_SIZE]) {
  const unsigned char *b = reinterpret_cast<const unsigned char *>(u.bytes);
  snprintf(out, GSB_UUID_BUFFER_SIZE,
           "GPU-%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", b[0], b[1], b[2],
           b[3], b[4], b[5], b[6], b[7], b[8], b[9], b[10], b[11], b[12], b[13], b[14], b[15]);
}

int load_libraries() {
  if (!G.cu.lib) {
    G.cu.lib = dlopen("libcuda.so.1", RTLD_LAZY | RTLD_GLOBAL);
    if (!G.cu.lib) {
      set_error("could not load CUDA driver library: %s", dlerror());
      return GSB_ERR_LIBRARY_NOT_FOUND;
    }
    DriverApi &c = G.cu;
    bool ok = load_sym(c.lib, c.cuInit, "cuInit") && load_sym(c.lib, c.cuDeviceGet, "cuDeviceGet") &&
              load_sym(c.lib, c.cuDeviceGetUuid, "cuDeviceGetUuid_v2", "cuDeviceGetUuid") &&
              load_sym(c.lib, c.cuDeviceTotalMem, "cuDeviceTotalMem_v2") &&
              load_sym(c.lib, c.cuMemGetInfo, "cuMemGetInfo_v2") &&
              load_sym(c.lib, c.cuMemGetAllocationGranularity, "cuMemGetAllocationGranularity") &&
              load_sym(c.lib, c.cuMemAddressReserve, "cuMemAddressReserve") &&
              load_sym(c.lib, c.cuMemAddressFree, "cuMemAddressFree") &&
              load_sym(c.lib, c.cuMemCreate, "cuMemCreate") && load_sym(c.lib, c.cuMemRelease, "cuMemRelease") &&
              load_sym(c.lib, c.cuMemMap, "cuMemMap") && load_sym(c.lib, c.cuMemUnmap, "cuMemUnmap") &&
              load_sym(c.lib, c.cuMemSetAccess, "cuMemSetAccess") &&
              load_sym(c.lib, c.cuGetErrorString, "cuGetErrorString");
    if (!ok) {
      set_error("libcuda.so.1 lacks a required entry point");
      return GSB_ERR_LIBRARY_NOT_FOUND;
    }
  }
  if (!G.ml.lib) {
    // same library name and flags as the reference's shim (nvml_dl.c:23)
    G.ml.lib = dlopen("libnvidia-ml.so.1", RTLD_LAZY | RTLD_GLOBAL);
    if (!G.ml.lib) {
      set_error("could not load NVML library");  // bindings.go:63
      return GSB_ERR_LIBRARY_NOT_FOUND;
    }
    NvmlApi &m = G.ml;
    bool ok = load_sym(m.lib, m.init, "nvmlInit_v2") && load_sym(m.lib, m.shutdown, "nvmlShutdown") &&
              load_sym(m.lib, m.errorString, "nvmlErrorString") &&
              load_sym(m.lib, m.getCount, "nvmlDeviceGetCount_v2") &&
              load_sym(m.lib, m.handleByIndex, "nvmlDeviceGetHandleByIndex_v2") &&
              load_sym(m.lib, m.getUUID, "nvmlDeviceGetUUID") &&
              load_sym(m.lib, m.getMinor, "nvmlDeviceGetMinorNumber") &&
              load_sym(m.lib, m.getMemoryInfo, "nvmlDeviceGetMemoryInfo") &&
              load_sym(m.lib, m.getPciInfo, "nvmlDeviceGetPciInfo_v3");
    if (!ok) {
      set_error("libnvidia-ml.so.1 lacks a required entry point");
      return GSB_ERR_LIBRARY_NOT_FOUND;
    }
    // optional (health): absence only disables the XID half
    load_sym(m.lib, m.eventSetCreate, "nvmlEventSetCreate");
    load_sym(m.lib, m.registerEvents, "nvmlDeviceRegisterEvents");
    load_sym(m.lib, m.eventSetWait, "nvmlEventSetWait_v2", "nvmlEventSetWait");
    load_sym(m.lib, m.eventSetFree, "nvmlEventSetFree");
  }
  return GSB_OK;
}

Device *device_at(uint32_t idx) {
  if (!G.inited) {
    set_error("gsb_init has not succeeded");
    return nullptr;
  }
  if (idx >= G.devs.size()) {
    set_error("nvml: device not found (index %u of %zu)", idx, G.devs.size());
    return nullptr;
  }
  return G.devs[idx].get();
}

// per-device CUDA resources, created on first probe/arena use (primary context via the runtime)
int ensure_ready(Device *d) {
  if (d->ordinal < 0) {
    set_error("device %s is not visible to the CUDA driver in this process", d->uuid);
    return GSB_ERR_NO_DEVICE;
  }
  if (d->cc_major != 10) {
    set_error("device %s is compute capability %u.%u; this library carries sm_100a code only", d->uuid,
              d->cc_major, d->cc_minor);
    return GSB_ERR_UNSUPPORTED_ARCH;
  }
  RT_TRY(cudaSetDevice(d->ordinal));
  if (d->ready) return GSB_OK;
  RT_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  RT_TRY(cudaEventCreate(&d->ev0));
  RT_TRY(cudaEventCreate(&d->ev1));
  RT_TRY(cudaHostAlloc(reinterpret_cast<void **>(&d->out_host), sizeof(gsb_kernel_out),
                       cudaHostAllocMapped | cudaHostAllocPortable));
  memset(d->out_host, 0, sizeof(gsb_kernel_out));
  RT_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d->out_dev), d->out_host, 0));
  RT_TRY(cudaMalloc(reinterpret_cast<void **>(&d->partials),
                    sizeof(gsb_partial) * gsb_kernel_max_grid((int)d->sm_count)));
  RT_TRY(cudaMalloc(reinterpret_cast<void **>(&d->ticket), sizeof(unsigned int)));
  RT_TRY(cudaMemset(d->ticket, 0, sizeof(unsigned int)));
  RT_TRY(cudaMalloc(reinterpret_cast<void **>(&d->tile_counter), sizeof(unsigned long long)));
  RT_TRY(cudaMemset(d->tile_counter, 0, sizeof(unsigned long long)));
  d->ready = true;
  return GSB_OK;
}

int arena_destroy_locked(Device *d) {
  if (!d->va) return GSB_OK;
  cudaSetDevice(d->ordinal);
  cudaStreamSynchronize(d->stream);
  size_t off = 0;
  for (const Chunk &c : d->chunks) {
    G.cu.cuMemUnmap(d->va + off, c.size);
    G.cu.cuMemRelease(c.handle);
    off += c.size;
  }
  d->chunks.clear();
  G.cu.cuMemAddressFree(d->va, d->va_size);
  d->va = 0;
  d->va_size = 0;
  d->arena_bytes = 0;
  if (d->seed_table) cudaFree(d->seed_table);
  d->seed_table = nullptr;
  d->gen.clear();
  return GSB_OK;
}

int run_probe_locked(Device *d, const gsb_probe_cfg *cfg, gsb_probe_result *out);

int arena_create_locked(Device *d, uint64_t max_bytes, uint64_t keep_free, uint64_t *arena_bytes) {
  int rc = ensure_ready(d);
  if (rc) return rc;
  arena_destroy_locked(d);

  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = d->ordinal;
  size_t gran = 0;
  CU_TRY(G.cu.cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  size_t free_b = 0, total_b = 0;
  CU_TRY(G.cu.cuMemGetInfo(&free_b, &total_b));
  uint64_t target = free_b > keep_free ? free_b - keep_free : 0;
  if (max_bytes && max_bytes < target) target = max_bytes;
  target = target / gran * gran;
  if (target == 0) {
    set_error("nothing allocatable on %s (free %zu, keep_free %llu)", d->uuid, free_b,
              (unsigned long long)keep_free);
    return GSB_ERR_OUT_OF_MEMORY;
  }
  CU_TRY(G.cu.cuMemAddressReserve(&d->va, target, 0, 0, 0));
  d->va_size = target;

  // Map physical chunks, largest first, until the driver refuses: what got mapped IS the
  // "actually allocatable" figure.
  const size_t ladder[] = {8ull << 30, 1ull << 30, 128ull << 20, 16ull << 20, 2ull << 20};
  uint64_t mapped = 0;
  for (size_t want : ladder) {
    size_t sz = want / gran * gran;
    if (sz == 0) sz = gran;
    while (mapped + sz <= target) {
      CUmemGenericAllocationHandle h;
      CUresult r = G.cu.cuMemCreate(&h, sz, &prop, 0);
      if (r != CUDA_SUCCESS) break;  // try the next smaller rung
      r = G.cu.cuMemMap(d->va + mapped, sz, 0, h, 0);
      if (r != CUDA_SUCCESS) {
        G.cu.cuMemRelease(h);
        set_error("cuMemMap -> %d (%s)", (int)r, cu_err(r));
        arena_destroy_locked(d);
        return GSB_ERR_DRIVER;
      }
      d->chunks.push_back({h, sz});
      mapped += sz;
    }
  }
  if (mapped == 0) {
    arena_destroy_locked(d);
    set_error("driver refused every allocation on %s", d->uuid);
    return GSB_ERR_OUT_OF_MEMORY;
  }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CUresult r = G.cu.cuMemSetAccess(d->va, mapped, &acc, 1);
  if (r != CUDA_SUCCESS) {
    set_error("cuMemSetAccess -> %d (%s)", (int)r, cu_err(r));
    arena_destroy_locked(d);
    return GSB_ERR_DRIVER;
  }
  d->arena_bytes = mapped;
  const size_t n_gran = (size_t)((mapped + kGranuleBytes - 1) / kGranuleBytes);
  d->gen.assign(n_gran, 0u);
  if (cudaMalloc(reinterpret_cast<void **>(&d->seed_table), n_gran * sizeof(uint32_t)) != cudaSuccess) {
    arena_destroy_locked(d);
    set_error("cudaMalloc(seed table) failed");
    return GSB_ERR_OUT_OF_MEMORY;
  }
  cudaMemset(d->seed_table, 0, n_gran * sizeof(uint32_t));

  // First generation: write every word once (proves every mapped byte is writable) so that all
  // later cycles are VERIFY_REFILL.
  gsb_probe_cfg fill;
  memset(&fill, 0, sizeof fill);
  fill.op = GSB_OP_FILL;
  fill.variant = GSB_VARIANT_AUTO;
  fill.seed_write = d->next_gen;
  fill.flags = GSB_PROBE_SEED_TABLE;
  gsb_probe_result res;
  rc = run_probe_locked(d, &fill, &res);
  if (rc) {
    arena_destroy_locked(d);
    return rc;
  }
  d->next_gen++;
  if (arena_bytes) *arena_bytes = mapped;
  return GSB_OK;
}

struct ProbeFlight {
  gsb_probe_cfg cfg;
  gsb_launch_geom geom;
  gsb_kernel_args args;
  uint64_t bytes = 0;
  uint64_t t_begin = 0;
  bool timed = false;
  bool launched = false;
};

// validate + enqueue (does not wait)
int probe_begin_locked(Device *d, const gsb_probe_cfg *cfg, gsb_probe_result *out, ProbeFlight *fl) {
  fl->t_begin = now_ns();
  fl->cfg = *cfg;
  memset(out, 0, sizeof *out);
  out->first_bad_offset = UINT64_MAX;
  int rc = ensure_ready(d);
  if (rc) return out->status = rc;
  if (!d->va) {
    set_error("no arena on %s: call gsb_arena_create first", d->uuid);
    return out->status = GSB_ERR_NO_ARENA;
  }
  if (cfg->op < GSB_OP_FILL || cfg->op > GSB_OP_VERIFY_REFILL || (cfg->window_offset & 15) ||
      (cfg->window_bytes & 15) || cfg->window_offset > d->arena_bytes) {
    set_error("invalid probe window/op (op %u, offset %llu, bytes %llu)", cfg->op,
              (unsigned long long)cfg->window_offset, (unsigned long long)cfg->window_bytes);
    return out->status = GSB_ERR_INVALID_ARGUMENT;
  }
  const uint64_t bytes = cfg->window_bytes ? cfg->window_bytes : d->arena_bytes - cfg->window_offset;
  if (cfg->window_offset + bytes > d->arena_bytes) {
    set_error("probe window [%llu, +%llu) exceeds arena of %llu bytes", (unsigned long long)cfg->window_offset,
              (unsigned long long)bytes, (unsigned long long)d->arena_bytes);
    return out->status = GSB_ERR_INVALID_ARGUMENT;
  }
  const bool table = (cfg->flags & GSB_PROBE_SEED_TABLE) != 0;
  if (table && (cfg->window_offset & 0xFFFF)) {
    set_error("seed-table probes need a 64 KiB aligned window offset");
    return out->status = GSB_ERR_INVALID_ARGUMENT;
  }
  // occupancy + function attributes are queried once per (op, variant, grid) and device, not per launch
  const uint64_t gkey = ((uint64_t)cfg->op << 48) | ((uint64_t)cfg->variant << 32) | cfg->grid_ctas;
  int e = 0;
  auto git = d->geoms.find(gkey);
  if (git != d->geoms.end()) {
    fl->geom = git->second;
  } else {
    e = gsb_kernel_geometry(cfg->op, cfg->variant, cfg->grid_ctas, (int)d->sm_count, &fl->geom);
    if (!e) d->geoms.emplace(gkey, fl->geom);
  }
  if (e) {
    set_error("kernel geometry: %s", cudaGetErrorString((cudaError_t)e));
    return out->status = (e == (int)cudaErrorInvalidDeviceFunction || e == (int)cudaErrorNoKernelImageForDevice)
                             ? GSB_ERR_UNSUPPORTED_ARCH
                             : GSB_ERR_DRIVER;
  }
  gsb_kernel_args &a = fl->args;
  memset(&a, 0, sizeof a);
  a.base = reinterpret_cast<uint4 *>(d->va);
  a.first_word = cfg->window_offset >> 4;
  a.n_words = bytes >> 4;
  a.seed_expect = cfg->seed_expect;
  a.seed_write = cfg->seed_write;
  a.seed_table = table ? d->seed_table : nullptr;
  a.table_update = (table && cfg->op != GSB_OP_VERIFY) ? d->seed_table : nullptr;
  a.arena_words = d->arena_bytes >> 4;
  a.granule_shift = kGranuleShiftWords;
  a.launch_seq = ++d->launch_seq;
  a.partials = d->partials;
  a.ticket = d->ticket;
  a.tile_counter = d->tile_counter;
  a.out = d->out_dev;
  fl->bytes = bytes;
  fl->timed = (cfg->flags & GSB_PROBE_TIMED) != 0;
  if (bytes > 0) {
    if (fl->timed) RT_TRY(cudaEventRecord(d->ev0, d->stream));
    e = gsb_kernel_launch(cfg->op, &fl->geom, &a, d->stream);
    if (e) {
      set_error("kernel launch: %s", cudaGetErrorString((cudaError_t)e));
      return out->status = GSB_ERR_DRIVER;
    }
    if (fl->timed) RT_TRY(cudaEventRecord(d->ev1, d->stream));
    fl->launched = true;
  }
  return GSB_OK;
}

// wait for the launch and collect what the last CTA wrote into pinned host memory
const uint64_t kWatchdogNs = [] {
  const char *e = getenv("GSB_PROBE_WATCHDOG_MS");
  return (uint64_t)(e ? atoll(e) : 0) * 1000000ull;
}();

int probe_end_locked(Device *d, ProbeFlight *fl, gsb_probe_result *out) {
  const gsb_probe_cfg *cfg = &fl->cfg;
  const gsb_kernel_args &a = fl->args;
  if (fl->launched) {
    // The last CTA stores the launch sequence number into pinned host memory after the results
    // (__threadfence_system between them): poll that word for up to 2 ms before falling back to a
    // blocking wait — a sleeping cudaStreamSynchronize costs 10-30 us of wake-up latency, a tenth of
    // a 1 GiB-window cycle. The stream sync below then returns immediately and orders the events.
    {
      const volatile uint32_t *flag = &d->out_host->done_flag;
      const uint64_t deadline = now_ns() + 2000000ull;
      while (*flag != a.launch_seq) {
        for (int i = 0; i < 64; i++) __builtin_ia32_pause();
        if (now_ns() > deadline) break;
      }
    }
    // Optional completion watchdog (knob GSB_PROBE_WATCHDOG_MS, default 0 = off until it has been verified on a
    // GPU): without it a kernel that never finishes parks this thread in cudaStreamSynchronize for ever, so a
    // wedged device that raises no XID would stay Healthy. With it the wait polls the stream against a deadline
    // and the probe fails, which the prober reports like any other probe fault.
    if (kWatchdogNs && d->out_host->done_flag != a.launch_seq) {
      const uint64_t deadline = now_ns() + kWatchdogNs + fl->bytes / 10;  // + 1 ns per 10 B: 10 GB/s is "wedged"
      while (cudaStreamQuery(d->stream) == cudaErrorNotReady) {
        if (now_ns() > deadline) {
          set_error("probe kernel did not finish within %llu ms: device wedged?",
                    (unsigned long long)((kWatchdogNs + fl->bytes / 10) / 1000000ull));
          return out->status = GSB_ERR_DRIVER;
        }
        struct timespec ts = {0, 50000};
        nanosleep(&ts, nullptr);
      }
    }
    cudaError_t se = cudaStreamSynchronize(d->stream);
    if (se != cudaSuccess) {
      set_error("probe kernel failed: %s", cudaGetErrorString(se));
      return out->status = GSB_ERR_DRIVER;
    }
    gsb_kernel_out ko;
    memcpy(&ko, d->out_host, sizeof ko);  // written by the last CTA; the stream sync above orders it
    if (ko.done_flag != a.launch_seq || ko.words_done != a.n_words) {
      set_error("probe kernel did not complete the window (seq %u/%u, words %llu/%llu)", ko.done_flag,
                a.launch_seq, ko.words_done, a.n_words);
      return out->status = GSB_ERR_DRIVER;
    }
    out->mismatch_words = ko.mismatch_words;
    out->mismatch_bits = ko.mismatch_bits;
    out->first_bad_offset = ko.first_bad_word == ~0ull ? UINT64_MAX : ko.first_bad_word << 4;
    out->checksum_xor = ko.checksum_xor;
    out->checksum_sum = ko.checksum_sum;
    if (fl->timed) {
      float ms = 0.f;
      RT_TRY(cudaEventElapsedTime(&ms, d->ev0, d->ev1));
      out->kernel_ns = (uint64_t)((double)ms * 1e6);
    }
    if (a.table_update) {  // host mirror of what the kernel's last CTA wrote
      const uint64_t g0 = (cfg->window_offset + kGranuleBytes - 1) / kGranuleBytes;
      const uint64_t g1 = (cfg->window_offset + fl->bytes) / kGranuleBytes;
      const bool tail = cfg->window_offset + fl->bytes == d->arena_bytes && (d->arena_bytes % kGranuleBytes);
      for (uint64_t g = g0; g < g1 + (tail ? 1 : 0) && g < d->gen.size(); g++) d->gen[g] = cfg->seed_write;
    }
  }
  out->variant = fl->geom.variant;
  out->grid_ctas = fl->geom.grid;
  out->block_threads = fl->geom.block;
  out->bytes_walked = fl->bytes;
  out->bytes_read = cfg->op == GSB_OP_FILL ? 0 : fl->bytes;
  out->bytes_written = cfg->op == GSB_OP_VERIFY ? 0 : fl->bytes;
  out->wall_ns = now_ns() - fl->t_begin;
  return out->status = GSB_OK;
}

int run_probe_locked(Device *d, const gsb_probe_cfg *cfg, gsb_probe_result *out) {
  ProbeFlight fl;
  int rc = probe_begin_locked(d, cfg, out, &fl);
  if (rc) return rc;
  return probe_end_locked(d, &fl, out);
}

// diagnostic knob GSB_NVML_SERIAL=1: one NVML/driver query at a time across this process's device threads (the
// open question behind profiles/node_cycle_8gpu_bimodal_r01.txt: do concurrent queries convoy on the driver's lock?)
std::mutex g_nvml_serial_mu;
const bool kNvmlSerial = [] {
  const char *e = getenv("GSB_NVML_SERIAL");
  return e && atoi(e) != 0;
}();

int query_info(Device *d, gsb_device_info *out) {
  memset(out, 0, sizeof *out);
  std::unique_lock<std::mutex> serial(g_nvml_serial_mu, std::defer_lock);
  if (kNvmlSerial) serial.lock();
  // identity: re-read from both sides, every call
  char uuid[GSB_UUID_BUFFER_SIZE] = {0};
  ML_TRY(G.ml.getUUID(d->nvml, uuid, GSB_UUID_BUFFER_SIZE));
  unsigned minor = 0;
  ML_TRY(G.ml.getMinor(d->nvml, &minor));
  nvmlMemory_t mem;
  memset(&mem, 0, sizeof mem);
  ML_TRY(G.ml.getMemoryInfo(d->nvml, &mem));
  if (d->ordinal >= 0) {
    CUuuid cu;
    CU_TRY(G.cu.cuDeviceGetUuid(&cu, d->cudev));
    char cuda_uuid[GSB_UUID_BUFFER_SIZE];
    format_uuid(cu, cuda_uuid);
    if (strcmp(cuda_uuid, uuid) != 0) {
      set_error("identity mismatch on minor %u: CUDA %s vs NVML %s", d->minor, cuda_uuid, uuid);
      return GSB_ERR_IDENTITY_MISMATCH;
    }
  }
  snprintf(out->uuid, sizeof out->uuid, "%s", uuid);
  snprintf(out->bus_id, sizeof out->bus_id, "%s", d->bus_id);
  out->minor = minor;
  out->cuda_ordinal = d->ordinal;
  out->sm_count = d->sm_count;
  out->cc_major = d->cc_major;
  out->cc_minor = d->cc_minor;
  out->total_bytes = mem.total;
  out->total_mib = mem.total / (1024ull * 1024ull);  // bindings.go:346-349
  out->free_bytes = mem.free;
  out->cuda_total_bytes = d->cuda_total;
  return GSB_OK;
}

// run `fn` on the device's own persistent thread (created on first use). Hand-off is spin-then-block in
// both directions: a node cycle every few hundred microseconds never pays a futex wake-up (tens of us per
// hop, a tenth of a 1 GiB-window cycle), an idle daemon's workers park after ~200 us.
// knob GSB_WORKER_SPIN_US (default 200; 0 = pure condition-variable hand-off)
const uint64_t kSpinNs = [] {
  const char *e = getenv("GSB_WORKER_SPIN_US");
  return (uint64_t)(e ? atoi(e) : 200) * 1000ull;
}();

void worker_submit(Device *d, std::function<void()> fn) {
  std::unique_lock<std::mutex> lk(d->wmu);
  if (!d->worker.joinable()) {
    d->worker = std::thread([d] {
      for (;;) {
        // spin for a job first
        const uint64_t until = now_ns() + kSpinNs;
        while (!d->job_flag.load(std::memory_order_acquire) && !d->quit_flag.load(std::memory_order_acquire) &&
               now_ns() < until)
          __builtin_ia32_pause();
        std::unique_lock<std::mutex> wl(d->wmu);
        d->wcv.wait(wl, [d] { return d->job_ready || d->worker_quit; });
        if (d->worker_quit) return;
        d->job_ready = false;
        d->job_flag.store(false, std::memory_order_relaxed);
        std::function<void()> f = std::move(d->job);
        wl.unlock();
        f();
        wl.lock();
        d->job_done = true;
        d->done_flag.store(true, std::memory_order_release);
        d->wcv.notify_all();
      }
    });
  }
  d->job = std::move(fn);
  d->job_done = false;
  d->done_flag.store(false, std::memory_order_relaxed);
  d->job_ready = true;
  d->job_flag.store(true, std::memory_order_release);
  d->wcv.notify_all();
}

void worker_wait(Device *d) {
  const uint64_t until = now_ns() + 5 * kSpinNs;
  while (!d->done_flag.load(std::memory_order_acquire) && now_ns() < until) __builtin_ia32_pause();
  std::unique_lock<std::mutex> lk(d->wmu);
  d->wcv.wait(lk, [d] { return d->job_done; });
}

void worker_stop(Device *d) {
  {
    std::lock_guard<std::mutex> lk(d->wmu);
    d->worker_quit = true;
    d->quit_flag.store(true, std::memory_order_release);
  }
  d->wcv.notify_all();
  if (d->worker.joinable()) d->worker.join();
}

void push_event(const gsb_event &ev) {
  {
    std::lock_guard<std::mutex> lk(G.hmu);
    G.events.push_back(ev);
  }
  G.hcv.notify_all();
}

}  // namespace

// ====================================================================== C ABI

extern "C" {

int gsb_abi_version(void) { return (int)GSB_ABI_VERSION; }

const char *gsb_strerror(int status) {
  switch (status) {
    case GSB_OK: return "ok";
    case GSB_ERR_NOT_INITIALIZED: return "not initialized";
    case GSB_ERR_INVALID_ARGUMENT: return "invalid argument";
    case GSB_ERR_LIBRARY_NOT_FOUND: return "could not load NVML library";
    case GSB_ERR_DRIVER: return "CUDA driver error";
    case GSB_ERR_NVML: return "NVML error";
    case GSB_ERR_NO_DEVICE: return "nvml: device not found";
    case GSB_ERR_IDENTITY_MISMATCH: return "CUDA/NVML device identity mismatch";
    case GSB_ERR_BUFFER_TOO_SMALL: return "buffer too small";
    case GSB_ERR_OUT_OF_MEMORY: return "out of device memory";
    case GSB_ERR_UNSUPPORTED_ARCH: return "device is not sm_100: no kernel image, no fallback";
    case GSB_ERR_TIMEOUT: return "timeout";
    case GSB_ERR_NO_ARENA: return "no probe arena";
    case GSB_ERR_MALFORMED: return "malformed protobuf";
    case GSB_ERR_STOPPED: return "health queue stopped";
    default: return "unknown gsb status";
  }
}

int gsb_last_error(char *buf, size_t cap) {
  if (!buf || cap == 0) return GSB_ERR_INVALID_ARGUMENT;
  snprintf(buf, cap, "%s", tl_error);
  return GSB_OK;
}

int gsb_init(void) {
  std::unique_lock<std::shared_mutex> api_lk(G.api_mu);
  std::lock_guard<std::mutex> lk(G.mu);
  if (G.inited) return GSB_OK;
  int rc = load_libraries();
  if (rc) return rc;
  ML_TRY(G.ml.init());
  CU_TRY(G.cu.cuInit(0));

  unsigned n_ml = 0;
  ML_TRY(G.ml.getCount(&n_ml));
  int n_cu = 0;
  RT_TRY(cudaGetDeviceCount(&n_cu));
  struct CudaDev {
    CUdevice dev;
    char uuid[GSB_UUID_BUFFER_SIZE];
  };
  std::vector<CudaDev> cds((size_t)n_cu);
  for (int i = 0; i < n_cu; i++) {
    CU_TRY(G.cu.cuDeviceGet(&cds[i].dev, i));
    CUuuid u;
    CU_TRY(G.cu.cuDeviceGetUuid(&u, cds[i].dev));
    format_uuid(u, cds[i].uuid);
  }
  G.devs.clear();
  for (unsigned i = 0; i < n_ml; i++) {  // NVML index order = the reference's order (nvidia.go:59)
    auto d = std::make_unique<Device>();
    ML_TRY(G.ml.handleByIndex(i, &d->nvml));
    ML_TRY(G.ml.getUUID(d->nvml, d->uuid, GSB_UUID_BUFFER_SIZE));
    unsigned minor = 0;
    ML_TRY(G.ml.getMinor(d->nvml, &minor));
    d->minor = minor;
    nvmlPciInfo_t pci;
    memset(&pci, 0, sizeof pci);
    ML_TRY(G.ml.getPciInfo(d->nvml, &pci));
    snprintf(d->bus_id, sizeof d->bus_id, "%s", pci.busId);
    for (int c = 0; c < n_cu; c++) {
      if (strcmp(cds[c].uuid, d->uuid) == 0) {
        d->ordinal = c;
        d->cudev = cds[c].dev;
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, c);
        d->sm_count = (uint32_t)v;
        cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, c);
        d->cc_major = (uint32_t)v;
        cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, c);
        d->cc_minor = (uint32_t)v;
        size_t tot = 0;
        G.cu.cuDeviceTotalMem(&tot, d->cudev);
        d->cuda_total = tot;
        break;
      }
    }
    G.devs.push_back(std::move(d));
  }
  G.inited = true;
  return GSB_OK;
}

int gsb_shutdown(void) {
  gsb_health_stop();
  std::unique_lock<std::shared_mutex> api_lk(G.api_mu);  // waits for in-flight calls of other threads
  std::lock_guard<std::mutex> lk(G.mu);
  if (!G.inited) return GSB_OK;
  for (auto &d : G.devs) {
    worker_stop(d.get());
    std::lock_guard<std::mutex> dl(d->mu);
    arena_destroy_locked(d.get());
    if (d->ready) {
      cudaSetDevice(d->ordinal);
      cudaStreamDestroy(d->stream);
      cudaEventDestroy(d->ev0);
      cudaEventDestroy(d->ev1);
      cudaFreeHost(d->out_host);
      cudaFree(d->partials);
      cudaFree(d->ticket);
      cudaFree(d->tile_counter);
      d->ready = false;
    }
  }
  G.devs.clear();
  G.inited = false;
  nvmlReturn_t r = G.ml.shutdown();  // library handles stay open: re-init is cheap and safe
  if (r != NVML_SUCCESS) {
    set_error("nvml: %s", G.ml.errorString(r));
    return GSB_ERR_NVML;
  }
  return GSB_OK;
}

int gsb_device_count(uint32_t *n) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  if (!n) return GSB_ERR_INVALID_ARGUMENT;
  if (!G.inited) {
    set_error("gsb_init has not succeeded");
    return GSB_ERR_NOT_INITIALIZED;
  }
  unsigned c = 0;
  ML_TRY(G.ml.getCount(&c));  // live query, like nvml.GetDeviceCount
  *n = c < G.devs.size() ? c : (uint32_t)G.devs.size();
  return GSB_OK;
}

int gsb_device_info_get(uint32_t idx, gsb_device_info *out) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  if (!out) return GSB_ERR_INVALID_ARGUMENT;
  Device *d = device_at(idx);
  if (!d) return G.inited ? GSB_ERR_NO_DEVICE : GSB_ERR_NOT_INITIALIZED;
  int rc = query_info(d, out);
  out->index = idx;
  return rc;
}

int gsb_arena_create(uint32_t idx, uint64_t max_bytes, uint64_t keep_free_bytes, uint64_t *arena_bytes) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  Device *d = device_at(idx);
  if (!d) return G.inited ? GSB_ERR_NO_DEVICE : GSB_ERR_NOT_INITIALIZED;
  std::lock_guard<std::mutex> lk(d->mu);
  return arena_create_locked(d, max_bytes, keep_free_bytes, arena_bytes);
}

int gsb_arena_destroy(uint32_t idx) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  Device *d = device_at(idx);
  if (!d) return G.inited ? GSB_ERR_NO_DEVICE : GSB_ERR_NOT_INITIALIZED;
  std::lock_guard<std::mutex> lk(d->mu);
  return arena_destroy_locked(d);
}

int gsb_arena_bytes(uint32_t idx, uint64_t *arena_bytes) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  Device *d = device_at(idx);
  if (!d || !arena_bytes) return G.inited ? GSB_ERR_NO_DEVICE : GSB_ERR_NOT_INITIALIZED;
  std::lock_guard<std::mutex> lk(d->mu);
  *arena_bytes = d->arena_bytes;
  return d->va ? GSB_OK : GSB_ERR_NO_ARENA;
}

int gsb_probe(uint32_t idx, const gsb_probe_cfg *cfg, gsb_probe_result *out) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  if (!cfg || !out) return GSB_ERR_INVALID_ARGUMENT;
  Device *d = device_at(idx);
  if (!d) {
    memset(out, 0, sizeof *out);
    return out->status = G.inited ? GSB_ERR_NO_DEVICE : GSB_ERR_NOT_INITIALIZED;
  }
  std::lock_guard<std::mutex> lk(d->mu);
  return run_probe_locked(d, cfg, out);
}

int gsb_probe_all(uint32_t n, const uint32_t *idxs, const gsb_probe_cfg *cfg, gsb_probe_result *results) {
  if (!idxs || !cfg || !results) return GSB_ERR_INVALID_ARGUMENT;
  std::vector<Device *> ds(n, nullptr);
  for (uint32_t i = 0; i < n; i++) {
    ds[i] = device_at(idxs[i]);
    if (!ds[i]) {
      memset(&results[i], 0, sizeof results[i]);
      results[i].status = G.inited ? GSB_ERR_NO_DEVICE : GSB_ERR_NOT_INITIALIZED;
      continue;
    }
    const uint32_t idx = idxs[i];
    gsb_probe_result *out = &results[i];
    worker_submit(ds[i], [idx, cfg, out] { gsb_probe(idx, cfg, out); });
  }
  int rc = GSB_OK;
  for (uint32_t i = 0; i < n; i++) {
    if (ds[i]) worker_wait(ds[i]);
    if (results[i].status != GSB_OK) rc = results[i].status;
  }
  return rc;
}

int64_t gsb_cycle_all(uint32_t n, const uint32_t *idxs, uint64_t cycle_no, uint64_t window_bytes, int unit_gib,
                      uint32_t variant, uint8_t *lw_buf, size_t lw_cap, gsb_cycle_result *results) {
  if (!idxs || !results || n == 0 || n > GSB_MAX_DEVICES) return GSB_ERR_INVALID_ARGUMENT;
  std::vector<Device *> ds(n, nullptr);
  std::vector<int> rcs(n, GSB_OK);
  std::vector<std::vector<uint8_t>> scratch(n);
  for (uint32_t i = 0; i < n; i++) {
    ds[i] = device_at(idxs[i]);
    if (!ds[i]) return G.inited ? GSB_ERR_NO_DEVICE : GSB_ERR_NOT_INITIALIZED;
  }
  for (uint32_t i = 0; i < n; i++) {  // fan out: every device runs its own cycle on its own thread
    scratch[i].resize(1 << 16);
    const uint32_t idx = idxs[i];
    gsb_cycle_result *out = &results[i];
    int *rc = &rcs[i];
    std::vector<uint8_t> *buf = &scratch[i];
    worker_submit(ds[i], [=] {
      *rc = gsb_cycle(idx, cycle_no, window_bytes, unit_gib, variant, buf->data(), buf->size(), out);
    });
  }
  for (uint32_t i = 0; i < n; i++) worker_wait(ds[i]);
  // the one join: concatenate in index order, slices from the first device, health from each verdict
  const uint32_t slices = results[0].slices;
  std::vector<const char *> uuids(n);
  std::vector<uint8_t> bits(((size_t)n * slices + 7) / 8, 0);
  bool any_bad = false;
  int rc_all = GSB_OK;
  for (uint32_t i = 0; i < n; i++) {
    if (rcs[i] < 0 && rcs[i] != GSB_ERR_DRIVER) rc_all = rcs[i];  // a failed launch is a verdict, not an API error
    uuids[i] = results[i].info.uuid;
    if (!results[i].healthy) {
      any_bad = true;
      for (uint32_t j = 0; j < slices; j++) {
        const size_t b = (size_t)i * slices + j;
        bits[b >> 3] |= (uint8_t)(1u << (b & 7));
      }
    }
  }
  if (rc_all) return rc_all;
  return gsb_encode_list_and_watch(uuids.data(), n, slices, any_bad ? bits.data() : nullptr, lw_buf, lw_cap);
}

int gsb_arena_read(uint32_t idx, uint64_t offset, void *dst, uint64_t bytes) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  Device *d = device_at(idx);
  if (!d || !dst) return GSB_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(d->mu);
  if (!d->va) return GSB_ERR_NO_ARENA;
  if (offset + bytes > d->arena_bytes) return GSB_ERR_INVALID_ARGUMENT;
  RT_TRY(cudaSetDevice(d->ordinal));
  RT_TRY(cudaMemcpy(dst, reinterpret_cast<const void *>(d->va + offset), bytes, cudaMemcpyDeviceToHost));
  return GSB_OK;
}

int gsb_arena_write(uint32_t idx, uint64_t offset, const void *src, uint64_t bytes) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  Device *d = device_at(idx);
  if (!d || !src) return GSB_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lk(d->mu);
  if (!d->va) return GSB_ERR_NO_ARENA;
  if (offset + bytes > d->arena_bytes) return GSB_ERR_INVALID_ARGUMENT;
  RT_TRY(cudaSetDevice(d->ordinal));
  RT_TRY(cudaMemcpy(reinterpret_cast<void *>(d->va + offset), src, bytes, cudaMemcpyHostToDevice));
  return GSB_OK;
}

int gsb_cycle(uint32_t idx, uint64_t cycle_no, uint64_t window_bytes, int unit_gib, uint32_t variant,
              uint8_t *lw_buf, size_t lw_cap, gsb_cycle_result *out) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  if (!out) return GSB_ERR_INVALID_ARGUMENT;
  memset(out, 0, sizeof *out);
  Device *d = device_at(idx);
  if (!d) return G.inited ? GSB_ERR_NO_DEVICE : GSB_ERR_NOT_INITIALIZED;
  if (window_bytes % kGranuleBytes) {
    set_error("cycle window must be a multiple of 64 MiB (or 0 = whole arena)");
    return GSB_ERR_INVALID_ARGUMENT;
  }
  std::lock_guard<std::mutex> lk(d->mu);
  if (!d->va) {
    set_error("no arena on %s: call gsb_arena_create first", d->uuid);
    return GSB_ERR_NO_ARENA;
  }
  // 1. health: enqueue the VERIFY_REFILL of this cycle's window (asynchronous)
  gsb_probe_cfg cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.op = GSB_OP_VERIFY_REFILL;
  cfg.variant = variant;
  cfg.flags = GSB_PROBE_TIMED | GSB_PROBE_SEED_TABLE;
  if (window_bytes && window_bytes < d->arena_bytes) {
    const uint64_t n_win = d->arena_bytes / window_bytes;
    cfg.window_offset = (cycle_no % n_win) * window_bytes;
    cfg.window_bytes = window_bytes;
  }
  cfg.seed_write = d->next_gen++;
  if (d->next_gen == 0) d->next_gen = 1;
  // diagnostic knob GSB_CYCLE_ORDER (tools/cycle_breakdown.py): 0 = launch, then inventory while the kernel
  // runs (shipped); 1 = inventory first, then launch; 2 = no inventory; 3 = a plain 170 us host sleep in
  // place of the inventory (tells host-side overlap apart from driver-side interference)
  static const int order = [] {
    const char *e = getenv("GSB_CYCLE_ORDER");
    return e ? atoi(e) : 0;
  }();
  ProbeFlight fl;
  int prc = GSB_OK;
  if (order != 1) prc = probe_begin_locked(d, &cfg, &out->probe, &fl);

  // 2. inventory while the kernel walks HBM: fresh identity + memory info, slices, S fake devices,
  //    wire bytes (optimistically with the health this device had going into the cycle)
  const uint64_t t0 = now_ns();
  int rc = GSB_OK;
  if (order == 2 || order == 3) {
    if (order == 3) {
      const uint64_t until = now_ns() + 170000;
      while (now_ns() < until) __builtin_ia32_pause();
    }
    snprintf(out->info.uuid, sizeof out->info.uuid, "%s", d->uuid);
    out->info.total_mib = 183359;
  } else {
    rc = query_info(d, &out->info);
  }
  if (rc == GSB_OK) {
    out->info.index = idx;
    out->slices = gsb_slices(out->info.total_mib, unit_gib);
  }
  const char *uuids[1] = {out->info.uuid};
  std::vector<uint8_t> bits;
  auto encode = [&](bool unhealthy) -> int64_t {
    if (unhealthy) bits.assign((out->slices + 7) / 8, 0xFF);  // every fake device of a faulted GPU (nvidia.go:146-150)
    return gsb_encode_list_and_watch(uuids, 1, out->slices, unhealthy ? bits.data() : nullptr, lw_buf, lw_cap);
  };
  if (rc == GSB_OK) out->lw_len = encode(d->faulted);
  out->inventory_ns = now_ns() - t0;
  if (order == 1) prc = probe_begin_locked(d, &cfg, &out->probe, &fl);

  // 3. verdict
  if (prc == GSB_OK) prc = probe_end_locked(d, &fl, &out->probe);
  if (rc) return rc;
  if (out->lw_len < 0) return (int)out->lw_len;
  out->healthy = (prc == GSB_OK && out->probe.mismatch_words == 0) ? 1u : 0u;
  if (!out->healthy && !d->faulted) {
    d->faulted = true;  // sticky (server.go:180 FIXME): the list this cycle reports already carries it
    out->lw_len = encode(true);
    if (out->lw_len < 0) return (int)out->lw_len;
  }
  return prc;
}

// ---------------------------------------------------------------------- health

int gsb_xid_is_benign(uint64_t xid) { return xid == 31 || xid == 43 || xid == 45; }

int gsb_health_set_recovery(uint32_t clean_cycles) {
  G.recovery_cycles = clean_cycles;
  return GSB_OK;
}

int gsb_health_inject(const gsb_event *ev) {
  if (!ev) return GSB_ERR_INVALID_ARGUMENT;
  push_event(*ev);
  return GSB_OK;
}

int gsb_health_wait(uint32_t timeout_ms, gsb_event *ev) {
  if (!ev) return GSB_ERR_INVALID_ARGUMENT;
  std::unique_lock<std::mutex> lk(G.hmu);
  const uint64_t gen = G.stop_gen;
  G.hcv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return !G.events.empty() || G.stop_gen != gen; });
  if (!G.events.empty()) {
    *ev = G.events.front();
    G.events.pop_front();
    return GSB_OK;
  }
  return G.stop_gen != gen ? GSB_ERR_STOPPED : GSB_ERR_TIMEOUT;
}

int gsb_health_start(uint32_t probe_period_ms, uint64_t window_bytes) {
  std::shared_lock<std::shared_mutex> api_lk(G.api_mu);
  if (!G.inited) return GSB_ERR_NOT_INITIALIZED;
  std::lock_guard<std::mutex> lk(G.mu);
  if (G.health_running) return GSB_OK;
  G.health_stop = false;
  // XID half: one registration per GPU (the reference registers the same GPU once per fake device)
  if (G.ml.eventSetCreate && G.ml.registerEvents && G.ml.eventSetWait && G.ml.eventSetFree &&
      G.ml.eventSetCreate(&G.event_set) == NVML_SUCCESS) {
    G.have_event_set = true;
    for (auto &d : G.devs) {
      nvmlReturn_t r = G.ml.registerEvents(d->nvml, nvmlEventTypeXidCriticalError, G.event_set);
      if (r == NVML_ERROR_NOT_SUPPORTED) {
        // nvidia.go:107-112: too old to support health checking -> marked unhealthy
        gsb_event ev;
        memset(&ev, 0, sizeof ev);
        snprintf(ev.uuid, sizeof ev.uuid, "%s", d->uuid);
        ev.etype = GSB_EVENT_XID;
        ev.edata = ~0ull;
        push_event(ev);
      }
    }
    G.health_threads.emplace_back([] {
      while (!G.health_stop.load()) {
        nvmlEventData_t data;
        memset(&data, 0, sizeof data);
        nvmlReturn_t r = G.ml.eventSetWait(G.event_set, &data, 200);
        if (r != NVML_SUCCESS) {
          if (r != NVML_ERROR_TIMEOUT) std::this_thread::sleep_for(std::chrono::milliseconds(200));
          continue;
        }
        if (data.eventType != nvmlEventTypeXidCriticalError) continue;  // nvidia.go:127-129
        gsb_event ev;
        memset(&ev, 0, sizeof ev);
        if (data.device) G.ml.getUUID(data.device, ev.uuid, GSB_UUID_BUFFER_SIZE);
        ev.etype = data.eventType;
        ev.edata = data.eventData;
        push_event(ev);
      }
    });
  }
  // active half: rotate a window probe over each device's arena
  if (probe_period_ms > 0) {
    for (uint32_t i = 0; i < G.devs.size(); i++) {
      G.health_threads.emplace_back([i, probe_period_ms, window_bytes] {
        std::vector<uint8_t> buf(1 << 16);
        uint64_t cycle = 0;
        bool reported = false;
        uint32_t clean = 0;
        while (!G.health_stop.load()) {
          gsb_cycle_result cr;
          int rc = gsb_cycle(i, cycle++, window_bytes, 1, GSB_VARIANT_AUTO, buf.data(), buf.size(), &cr);
          const bool this_cycle_clean = rc == GSB_OK && cr.probe.mismatch_words == 0;
          if (rc != GSB_ERR_NO_ARENA && !this_cycle_clean && !reported) {
            gsb_event ev;
            memset(&ev, 0, sizeof ev);
            snprintf(ev.uuid, sizeof ev.uuid, "%s", G.devs[i]->uuid);
            ev.etype = GSB_EVENT_PROBE;
            ev.edata = rc == GSB_OK ? GSB_PROBE_FAULT_MISMATCH : GSB_PROBE_FAULT_LAUNCH;
            push_event(ev);
            reported = true;
            clean = 0;
          } else if (reported) {
            clean = this_cycle_clean ? clean + 1 : 0;
            const uint32_t need = G.recovery_cycles.load();
            if (need > 0 && clean >= need) {  // optional recovery: the refills since the fault all verified
              {
                std::lock_guard<std::mutex> dl(G.devs[i]->mu);
                G.devs[i]->faulted = false;
              }
              gsb_event ev;
              memset(&ev, 0, sizeof ev);
              snprintf(ev.uuid, sizeof ev.uuid, "%s", G.devs[i]->uuid);
              ev.etype = GSB_EVENT_PROBE;
              ev.edata = GSB_PROBE_RECOVERED;
              push_event(ev);
              reported = false;
              clean = 0;
            }
          }
          for (uint32_t slept = 0; slept < probe_period_ms && !G.health_stop.load(); slept += 10)
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
      });
    }
  }
  G.health_running = true;
  return GSB_OK;
}

int gsb_health_stop(void) {
  {  // wake every waiter, whether or not the internal threads are running
    std::lock_guard<std::mutex> lk(G.hmu);
    G.stop_gen++;
  }
  G.hcv.notify_all();
  std::vector<std::thread> ts;
  {
    std::lock_guard<std::mutex> lk(G.mu);
    if (!G.health_running) return GSB_OK;
    G.health_stop = true;
    ts.swap(G.health_threads);
    G.health_running = false;
  }
  for (auto &t : ts) t.join();
  if (G.have_event_set) {
    G.ml.eventSetFree(G.event_set);
    G.have_event_set = false;
  }
  std::lock_guard<std::mutex> lk(G.hmu);
  G.events.clear();
  G.health_stop = false;
  return GSB_OK;
}

}  // extern "C"
