/*
 * ref_inventory.c — C restatement of the reference's inventory + health path, linked against the
 * reference's OWN NVML shim. TEST INFRASTRUCTURE / CPU BASELINE ONLY: never linked into or called
 * by the product; only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline, --impl
 * reference) may execute the binary built from it (oracle/_ref/ref_inventory).
 *
 * Why C and not the reference binary: the reference is Go 1.10 + cgo and there is no Go toolchain
 * in the build image. Its only native code — vendor/github.com/NVIDIA/
 * gpu-monitoring-tools/bindings/go/nvml/{nvml_dl.c,nvml_dl.h,nvml.h} — does compile, so this file
 * is compiled TOGETHER WITH that nvml_dl.c (from where it lies under /root/reference; see Makefile)
 * and with the reference's link flags (bindings.go:5), and restates the Go call sequence around it
 * call for call:
 *
 *   nvml.Init            nvml.go:250 -> bindings.go:60-66 -> nvml_dl.c:21-28
 *   getDeviceCount       pkg/gpu/nvidia/nvidia.go:47-51 -> bindings.go:166-171
 *   nvml.NewDevice       nvml.go:297-359 (11 getters, nil checks :327-329, numaNode :266-281)
 *   deviceGetMemoryInfo  bindings.go:333-364   (total /= 1024*1024)
 *   setGPUMemory         pkg/gpu/nvidia/nvidia.go:34-41
 *   getDevices fan-out   pkg/gpu/nvidia/nvidia.go:59-86 (Sscanf of the path :65, Sprintf of IDs :27)
 *   gogo marshal         vendor/k8s.io/kubernetes/pkg/kubelet/apis/deviceplugin/v1beta1/api.pb.go:794-843
 *   watchXIDs            pkg/gpu/nvidia/nvidia.go:100-152: NewEventSet (bindings.go:68-73),
 *                        RegisterEventForDevice per FAKE device (bindings.go:97-128: count + linear
 *                        HandleByIndex/GetUUID scan), WaitForEvent (bindings.go:134-146), DeleteEventSet
 *
 * Being C, it has none of Go's per-call cgo transition, allocation, or glog cost: as a CPU baseline
 * it flatters the reference.
 *
 * Usage:
 *   ref_inventory inventory [--unit GiB|MiB] [--lw-out FILE] [--gpus N]   one pass, JSON on stdout
 *   ref_inventory bench --iters K [--warmup W] [--wait-ms T] [--unit GiB] [--gpus N] [--setup-iters M]
 *       phases as the reference runs them: health SET-UP once per plugin start (NewEventSet + one
 *       RegisterEventForDevice per FAKE device, nvidia.go:101-117), then K timed cycles of
 *       inventory (getDevices + marshal) + one health poll (WaitForEvent on the standing event set). JSON on stdout.
 *   --gpus N: behave as on a node with only the first N GPUs (GetCount is capped to N everywhere it is asked).
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "nvml_dl.h" /* the reference's header (includes its vendored nvml.h, API v9) */

#define MAX_GPUS 64

static unsigned g_gpu_limit = 0; /* --gpus N: 0 = every GPU NVML reports */

/* nvmlDeviceGetCount as a node with --gpus N devices would answer it */
static nvmlReturn_t device_count(unsigned *n) {
  nvmlReturn_t r = nvmlDeviceGetCount(n);
  if (r == NVML_SUCCESS && g_gpu_limit && *n > g_gpu_limit) *n = g_gpu_limit;
  return r;
}

typedef struct {
  char uuid[NVML_DEVICE_UUID_BUFFER_SIZE];
  char path[64];
  char busid[NVML_DEVICE_PCI_BUS_ID_BUFFER_SIZE];
  char model[NVML_DEVICE_NAME_BUFFER_SIZE];
  unsigned long long total_bytes;
  unsigned long long memory_mib;
  unsigned minor;
  unsigned numa;
} ref_device;

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

static void die(const char *what, nvmlReturn_t r) {
  /* check(): log.Fatalln("Fatal:", err) with err = "nvml: <nvmlErrorString>" (nvidia.go:20-24) */
  fprintf(stderr, "Fatal: nvml: %s (%s)\n", r == NVML_ERROR_LIBRARY_NOT_FOUND ? "library not found" : nvmlErrorString(r), what);
  exit(r == NVML_ERROR_LIBRARY_NOT_FOUND ? 12 : 1);
}

/* numaNode (nvml.go:266-281) */
static unsigned numa_node(const char *busid) {
  char p[128], lower[64];
  size_t n = strlen(busid);
  if (n < 4) return 0;
  for (size_t i = 0; i + 4 <= n; i++) lower[i] = (char)tolower((unsigned char)busid[i + 4]);
  lower[n - 4] = 0;
  snprintf(p, sizeof p, "/sys/bus/pci/devices/%s/numa_node", lower);
  FILE *f = fopen(p, "r");
  if (!f) return 0;
  long node = 0;
  if (fscanf(f, "%ld", &node) != 1) node = 0;
  fclose(f);
  return node < 0 ? 0 : (unsigned)node;
}

/* nvml.NewDevice (nvml.go:297-359): all 11 getters in the reference's order */
static void new_device(unsigned idx, ref_device *out) {
  nvmlDevice_t h;
  nvmlReturn_t r;
  if ((r = nvmlDeviceGetHandleByIndex(idx, &h)) != NVML_SUCCESS) die("HandleByIndex", r);
  r = nvmlDeviceGetName(h, out->model, sizeof out->model);
  if (r != NVML_SUCCESS && r != NVML_ERROR_NOT_SUPPORTED) die("GetName", r);
  r = nvmlDeviceGetUUID(h, out->uuid, sizeof out->uuid);
  if (r != NVML_SUCCESS) die("GetUUID", r); /* nil uuid -> ErrUnsupportedGPU (:327-329) */
  r = nvmlDeviceGetMinorNumber(h, &out->minor);
  if (r != NVML_SUCCESS) die("GetMinorNumber", r);
  unsigned power = 0;
  r = nvmlDeviceGetPowerManagementLimit(h, &power);
  if (r != NVML_SUCCESS && r != NVML_ERROR_NOT_SUPPORTED) die("GetPowerManagementLimit", r);
  nvmlMemory_t mem;
  r = nvmlDeviceGetMemoryInfo(h, &mem);
  if (r != NVML_SUCCESS) die("GetMemoryInfo", r);
  out->total_bytes = mem.total;
  out->memory_mib = mem.total / (1024 * 1024); /* bindings.go:346-349 */
  nvmlPciInfo_t pci;
  r = nvmlDeviceGetPciInfo(h, &pci);
  if (r != NVML_SUCCESS) die("GetPciInfo", r);
  snprintf(out->busid, sizeof out->busid, "%s", pci.busId);
  nvmlBAR1Memory_t bar1;
  r = nvmlDeviceGetBAR1MemoryInfo(h, &bar1);
  if (r != NVML_SUCCESS && r != NVML_ERROR_NOT_SUPPORTED) die("GetBAR1MemoryInfo", r);
  unsigned gen = 0, width = 0, sm = 0, memclk = 0;
  r = nvmlDeviceGetMaxPcieLinkGeneration(h, &gen);
  if (r != NVML_SUCCESS && r != NVML_ERROR_NOT_SUPPORTED) die("GetMaxPcieLinkGeneration", r);
  r = nvmlDeviceGetMaxPcieLinkWidth(h, &width);
  if (r != NVML_SUCCESS && r != NVML_ERROR_NOT_SUPPORTED) die("GetMaxPcieLinkWidth", r);
  r = nvmlDeviceGetMaxClockInfo(h, NVML_CLOCK_SM, &sm);
  if (r == NVML_SUCCESS) r = nvmlDeviceGetMaxClockInfo(h, NVML_CLOCK_MEM, &memclk);
  if (r != NVML_SUCCESS && r != NVML_ERROR_NOT_SUPPORTED) die("GetMaxClockInfo", r);
  snprintf(out->path, sizeof out->path, "/dev/nvidia%u", out->minor); /* nvml.go:330 */
  out->numa = numa_node(out->busid);
}

typedef struct {
  char **ids; /* fake device IDs, heap strings like Go's */
  size_t n;
  unsigned gpu_memory; /* the process-global gpuMemory */
  unsigned minors[MAX_GPUS];
} ref_devs;

/* getDevices (nvidia.go:53-89) */
static void get_devices(int unit_gib, ref_device *devs, unsigned *n_out, ref_devs *out) {
  unsigned n = 0;
  nvmlReturn_t r = device_count(&n);
  if (r != NVML_SUCCESS) die("GetCount", r);
  if (n > MAX_GPUS) n = MAX_GPUS;
  out->ids = NULL;
  out->n = 0;
  out->gpu_memory = 0;
  size_t cap = 0;
  for (unsigned i = 0; i < n; i++) {
    new_device(i, &devs[i]);
    unsigned id = 0;
    if (sscanf(devs[i].path, "/dev/nvidia%u", &id) != 1) { /* nvidia.go:65 */
      fprintf(stderr, "Fatal: input does not match format\n");
      exit(1);
    }
    out->minors[i] = id;
    if (out->gpu_memory == 0) { /* nvidia.go:70-72 + setGPUMemory :34-41 */
      unsigned raw = (unsigned)devs[i].memory_mib;
      out->gpu_memory = unit_gib ? raw / 1024 : raw;
    }
    for (unsigned j = 0; j < out->gpu_memory; j++) { /* nvidia.go:73-85 */
      if (out->n == cap) {
        cap = cap ? cap * 2 : 256; /* append() growth */
        out->ids = (char **)realloc(out->ids, cap * sizeof(char *));
      }
      char *s = NULL;
      if (asprintf(&s, "%s-_-%u", devs[i].uuid, j) < 0) exit(1); /* fmt.Sprintf, one alloc per ID */
      out->ids[out->n++] = s;
    }
  }
  *n_out = n;
}

static void free_devs(ref_devs *d) {
  for (size_t i = 0; i < d->n; i++) free(d->ids[i]);
  free(d->ids);
  d->ids = NULL;
  d->n = 0;
}

static size_t sov(uint64_t x) {
  size_t n = 1;
  while (x >= 0x80) {
    x >>= 7;
    n++;
  }
  return n;
}
static uint8_t *put_varint(uint8_t *p, uint64_t v) {
  while (v >= 0x80) {
    *p++ = (uint8_t)(v | 0x80);
    v >>= 7;
  }
  *p++ = (uint8_t)v;
  return p;
}

/* ListAndWatchResponse.Marshal: Size() pass, then MarshalTo (api.pb.go:786-843) */
static uint8_t *marshal_lw(const ref_devs *d, size_t *len) {
  static const char healthy[] = "Healthy";
  size_t total = 0;
  for (size_t i = 0; i < d->n; i++) {
    size_t l = strlen(d->ids[i]);
    size_t dev = 1 + sov(l) + l + 1 + sov(7) + 7;
    total += 1 + sov(dev) + dev;
  }
  uint8_t *buf = (uint8_t *)malloc(total ? total : 1), *p = buf;
  for (size_t i = 0; i < d->n; i++) {
    size_t l = strlen(d->ids[i]);
    size_t dev = 1 + sov(l) + l + 1 + sov(7) + 7;
    *p++ = 0x0a;
    p = put_varint(p, dev);
    *p++ = 0x0a;
    p = put_varint(p, l);
    memcpy(p, d->ids[i], l);
    p += l;
    *p++ = 0x12;
    p = put_varint(p, 7);
    memcpy(p, healthy, 7);
    p += 7;
  }
  *len = (size_t)(p - buf);
  return buf;
}

/* RegisterEventForDevice (bindings.go:97-128) */
static nvmlReturn_t register_event_for_device(nvmlEventSet_t set, const char *uuid, unsigned long *calls) {
  unsigned n = 0;
  nvmlReturn_t r = device_count(&n);
  (*calls)++;
  if (r != NVML_SUCCESS) return r;
  for (unsigned i = 0; i < n; i++) {
    nvmlDevice_t h;
    r = nvmlDeviceGetHandleByIndex(i, &h);
    (*calls)++;
    if (r != NVML_SUCCESS) return r;
    char duuid[NVML_DEVICE_UUID_BUFFER_SIZE];
    r = nvmlDeviceGetUUID(h, duuid, sizeof duuid);
    (*calls)++;
    if (r != NVML_SUCCESS) return r;
    if (strcmp(duuid, uuid) != 0) continue;
    (*calls)++;
    return nvmlDeviceRegisterEvents(h, nvmlEventTypeXidCriticalError, set);
  }
  return NVML_ERROR_NOT_FOUND;
}

static uint64_t fnv1a(const uint8_t *p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

static int cmp_double(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return x < y ? -1 : x > y;
}
static double pct(double *v, int n, double q) {
  int i = (int)(q * (n - 1) + 0.5);
  return v[i < 0 ? 0 : i >= n ? n - 1 : i];
}

/* entry point of libref_inventory.so; oracle/ref_launcher.c dlopen()s it (see Makefile for why) */
int ref_main(int argc, char **argv) {
  const char *mode = argc > 1 ? argv[1] : "inventory";
  int unit_gib = 1, iters = 30, setup_iters = 1, warmup = 0;
  unsigned wait_ms = 0;
  const char *lw_out = NULL;
  for (int i = 2; i < argc; i++) {
    if (!strcmp(argv[i], "--unit") && i + 1 < argc) unit_gib = strcmp(argv[++i], "MiB") != 0;
    else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--wait-ms") && i + 1 < argc) wait_ms = (unsigned)atoi(argv[++i]);
    else if (!strcmp(argv[i], "--lw-out") && i + 1 < argc) lw_out = argv[++i];
    else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) g_gpu_limit = (unsigned)atoi(argv[++i]);
    else if (!strcmp(argv[i], "--setup-iters") && i + 1 < argc) setup_iters = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
  }

  double t0 = now_us();
  nvmlReturn_t r = nvmlInit_dl(); /* the reference's function, from the reference's nvml_dl.c */
  double init_us = now_us() - t0;
  if (r == NVML_ERROR_LIBRARY_NOT_FOUND) {
    fprintf(stderr, "could not load NVML library\n"); /* bindings.go:63 */
    return 12;
  }
  if (r != NVML_SUCCESS) die("Init", r);

  static ref_device devs[MAX_GPUS];
  ref_devs fan;
  unsigned n = 0;

  if (!strcmp(mode, "inventory")) {
    get_devices(unit_gib, devs, &n, &fan);
    size_t len = 0;
    uint8_t *lw = marshal_lw(&fan, &len);
    if (lw_out) {
      FILE *f = fopen(lw_out, "wb");
      if (!f || fwrite(lw, 1, len, f) != len) return 2;
      fclose(f);
    }
    printf("{\"mode\":\"inventory\",\"unit\":\"%s\",\"nvml_init_us\":%.1f,\"n_gpus\":%u,\"gpu_memory\":%u,"
           "\"n_devices\":%zu,\"lw_len\":%zu,\"lw_fnv1a\":\"%016llx\",\"devices\":[",
           unit_gib ? "GiB" : "MiB", init_us, n, fan.gpu_memory, fan.n, len, (unsigned long long)fnv1a(lw, len));
    for (unsigned i = 0; i < n; i++)
      printf("%s{\"index\":%u,\"uuid\":\"%s\",\"path\":\"%s\",\"minor\":%u,\"bus_id\":\"%s\",\"model\":\"%s\","
             "\"total_bytes\":%llu,\"memory_mib\":%llu,\"numa\":%u}",
             i ? "," : "", i, devs[i].uuid, devs[i].path, fan.minors[i], devs[i].busid, devs[i].model,
             devs[i].total_bytes, devs[i].memory_mib, devs[i].numa);
    printf("],\"first_id\":\"%s\",\"last_id\":\"%s\"}\n", fan.n ? fan.ids[0] : "", fan.n ? fan.ids[fan.n - 1] : "");
    free(lw);
    free_devs(&fan);
  } else if (!strcmp(mode, "bench")) {
    if (iters < 1) iters = 1;
    if (setup_iters < 1) setup_iters = 1;
    double *t_inv = malloc(sizeof(double) * iters), *t_wait = malloc(sizeof(double) * iters),
           *t_cyc = malloc(sizeof(double) * iters), *t_reg = malloc(sizeof(double) * setup_iters);
    unsigned long reg_calls = 0;
    int reg_rc = 0, wait_rc = 0;
    size_t n_dev = 0, lw_len = 0;
    /* ---- set-up, once per plugin start: the first getDevices (server.go:39) and watchXIDs' registration loop,
     *      once per FAKE device (nvidia.go:101-117). Repeated --setup-iters times only to get a distribution; the
     *      event set of the last repetition stays registered for the timed cycles. ---- */
    nvmlEventSet_t set;
    int have_set = 0;
    double first_inventory_us = 0;
    for (int si = 0; si < setup_iters; si++) {
      double a0 = now_us();
      get_devices(unit_gib, devs, &n, &fan);
      if (si == 0) first_inventory_us = now_us() - a0;
      if (have_set) nvmlEventSetFree(set);
      double b = now_us();
      nvmlEventSetCreate(&set);
      have_set = 1;
      reg_calls = 0;
      for (size_t i = 0; i < fan.n; i++) {
        char real[NVML_DEVICE_UUID_BUFFER_SIZE];
        const char *sep = strstr(fan.ids[i], "-_-"); /* extractRealDeviceID */
        size_t l = sep ? (size_t)(sep - fan.ids[i]) : strlen(fan.ids[i]);
        memcpy(real, fan.ids[i], l);
        real[l] = 0;
        nvmlReturn_t rr = register_event_for_device(set, real, &reg_calls);
        if (rr != NVML_SUCCESS) reg_rc = (int)rr;
      }
      t_reg[si] = now_us() - b;
      free_devs(&fan);
    }
    /* ---- steady state: K cycles of inventory + one health poll on the standing event set ---- */
    for (int it = -warmup; it < iters; it++) { /* it < 0: untimed warm-up cycles */
      double a = now_us();
      /* (i) inventory: getDevices + marshal of the ListAndWatch send */
      get_devices(unit_gib, devs, &n, &fan);
      size_t len = 0;
      uint8_t *lw = marshal_lw(&fan, &len);
      double b = now_us();
      /* (ii) one health poll: WaitForEvent(set, wait_ms) incl. its trailing GetUUID (bindings.go:134-146) */
      nvmlEventData_t data;
      memset(&data, 0, sizeof data);
      nvmlReturn_t wr = nvmlEventSetWait(set, &data, wait_ms);
      wait_rc = (int)wr;
      if (wr == NVML_SUCCESS) {
        char u[NVML_DEVICE_UUID_BUFFER_SIZE];
        nvmlDeviceGetUUID(data.device, u, sizeof u);
      }
      double d = now_us();
      if (it >= 0) {
        t_inv[it] = b - a;
        t_wait[it] = d - b;
        t_cyc[it] = d - a;
      }
      n_dev = fan.n;
      lw_len = len;
      free(lw);
      free_devs(&fan);
    }
    nvmlEventSetFree(set);
    double sum = 0, sum_inv = 0, sum_wait = 0;
    for (int i = 0; i < iters; i++) {
      sum += t_cyc[i];
      sum_inv += t_inv[i];
      sum_wait += t_wait[i];
    }
    qsort(t_inv, iters, sizeof(double), cmp_double);
    qsort(t_reg, setup_iters, sizeof(double), cmp_double);
    qsort(t_wait, iters, sizeof(double), cmp_double);
    qsort(t_cyc, iters, sizeof(double), cmp_double);
    printf("{\"mode\":\"bench\",\"iters\":%d,\"setup_iters\":%d,\"n_gpus\":%u,\"gpu_limit\":%u,\"n_devices\":%zu,"
           "\"lw_len\":%zu,\"wait_ms\":%u,"
           "\"nvml_init_us\":%.1f,\"first_inventory_us\":%.1f,\"register_calls_per_setup\":%lu,\"register_rc\":%d,\"wait_rc\":%d,"
           "\"inventory_us\":{\"p10\":%.1f,\"p50\":%.1f,\"p90\":%.1f,\"mean\":%.1f},"
           "\"health_setup_us\":{\"p10\":%.1f,\"p50\":%.1f,\"p90\":%.1f},"
           "\"health_poll_us\":{\"p10\":%.1f,\"p50\":%.1f,\"p90\":%.1f,\"mean\":%.1f},"
           "\"cycle_us\":{\"p10\":%.1f,\"p50\":%.1f,\"p90\":%.1f,\"p99\":%.1f,\"mean\":%.1f},\"total_us\":%.1f}\n",
           iters, setup_iters, n, g_gpu_limit, n_dev, lw_len, wait_ms, init_us, first_inventory_us, reg_calls, reg_rc, wait_rc,
           pct(t_inv, iters, .1), pct(t_inv, iters, .5), pct(t_inv, iters, .9), sum_inv / iters,
           pct(t_reg, setup_iters, .1), pct(t_reg, setup_iters, .5), pct(t_reg, setup_iters, .9),
           pct(t_wait, iters, .1), pct(t_wait, iters, .5), pct(t_wait, iters, .9), sum_wait / iters,
           pct(t_cyc, iters, .1), pct(t_cyc, iters, .5), pct(t_cyc, iters, .9), pct(t_cyc, iters, .99), sum / iters, sum);
  } else {
    fprintf(stderr, "unknown mode %s\n", mode);
    return 2;
  }
  nvmlShutdown_dl();
  return 0;
}
