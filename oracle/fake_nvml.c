/*
 * fake_nvml.c — a stand-in libnvidia-ml.so.1 with 8 synthetic B200s (test infrastructure only).
 * Lets oracle/_ref/ref_inventory (and therefore the reference's own nvml_dl.c) run on a GPU-less
 * box: LD_LIBRARY_PATH=oracle/_fake. Values mirror what the real box reported
 * (profiles/envprobe_r01.txt): total 192265846784 B, driver 580.159.03. FAKE_NVML_GPUS overrides
 * the count; FAKE_NVML_XID=<idx>:<xid> makes the next nvmlEventSetWait deliver that event.
 */
#include <nvml.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#undef nvmlInit
#undef nvmlDeviceGetCount
#undef nvmlDeviceGetHandleByIndex
#undef nvmlDeviceGetPciInfo
#undef nvmlEventSetWait

static int n_gpus(void) {
  const char *e = getenv("FAKE_NVML_GPUS");
  int n = e ? atoi(e) : 8;
  return n < 0 ? 0 : n > 16 ? 16 : n;
}
static const unsigned kBus[16] = {0x1b, 0x43, 0x52, 0x61, 0x9d, 0xc3, 0xd1, 0xdf, 0xe1, 0xe3, 0xe5, 0xe7, 0xe9, 0xeb, 0xed, 0xef};
static const unsigned kMinor[16] = {2, 3, 0, 1, 6, 7, 4, 5, 8, 9, 10, 11, 12, 13, 14, 15}; /* minor != index, as on real HGX boxes */
static unsigned long registered;
static long idx_of(nvmlDevice_t d) { return (long)(size_t)d - 1; }

nvmlReturn_t nvmlInit_v2(void) { return NVML_SUCCESS; }
nvmlReturn_t nvmlInitWithFlags(unsigned flags) { (void)flags; return NVML_SUCCESS; }  /* what pynvml.nvmlInit() calls */
nvmlReturn_t nvmlShutdown(void) { return NVML_SUCCESS; }
const char *nvmlErrorString(nvmlReturn_t r) { return r == NVML_SUCCESS ? "Success" : r == NVML_ERROR_TIMEOUT ? "Timeout" : "Unknown Error"; }
nvmlReturn_t nvmlSystemGetDriverVersion(char *v, unsigned n) { snprintf(v, n, "580.159.03"); return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetCount_v2(unsigned *n) { *n = (unsigned)n_gpus(); return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetHandleByIndex_v2(unsigned i, nvmlDevice_t *d) {
  if ((int)i >= n_gpus()) return NVML_ERROR_INVALID_ARGUMENT;
  *d = (nvmlDevice_t)(size_t)(i + 1);
  return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetName(nvmlDevice_t d, char *s, unsigned n) { snprintf(s, n, "NVIDIA B200"); return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetUUID(nvmlDevice_t d, char *s, unsigned n) {
  long i = idx_of(d);
  if (i < 0) return NVML_ERROR_INVALID_ARGUMENT;
  snprintf(s, n, "GPU-%08lx-4820-abfc-e83e-9431819757%02lx", 0xfef80890ul + (unsigned long)i, (unsigned long)i);
  return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetMinorNumber(nvmlDevice_t d, unsigned *m) { *m = kMinor[idx_of(d)]; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetPowerManagementLimit(nvmlDevice_t d, unsigned *p) { *p = 1000000; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetMemoryInfo(nvmlDevice_t d, nvmlMemory_t *m) {
  m->total = 192265846784ull; m->used = 762839040ull; m->free = m->total - m->used;
  return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetPciInfo_v3(nvmlDevice_t d, nvmlPciInfo_t *p) {
  memset(p, 0, sizeof *p);
  snprintf(p->busId, sizeof p->busId, "00000000:%02X:00.0", kBus[idx_of(d)]);
  p->bus = kBus[idx_of(d)];
  return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetBAR1MemoryInfo(nvmlDevice_t d, nvmlBAR1Memory_t *b) { b->bar1Total = 1ull << 38; b->bar1Used = 1 << 20; b->bar1Free = b->bar1Total - b->bar1Used; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetMaxPcieLinkGeneration(nvmlDevice_t d, unsigned *g) { *g = 5; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetMaxPcieLinkWidth(nvmlDevice_t d, unsigned *w) { *w = 16; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetMaxClockInfo(nvmlDevice_t d, nvmlClockType_t t, unsigned *c) { *c = t == NVML_CLOCK_SM ? 1965 : 3996; return NVML_SUCCESS; }
nvmlReturn_t nvmlEventSetCreate(nvmlEventSet_t *s) { *s = (nvmlEventSet_t)(size_t)0x5e7; registered = 0; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceRegisterEvents(nvmlDevice_t d, unsigned long long t, nvmlEventSet_t s) { registered++; return NVML_SUCCESS; }
nvmlReturn_t nvmlEventSetFree(nvmlEventSet_t s) { return NVML_SUCCESS; }
static nvmlReturn_t wait_impl(nvmlEventData_t *data) {
  const char *e = getenv("FAKE_NVML_XID");
  unsigned i = 0, xid = 0;
  if (e && sscanf(e, "%u:%u", &i, &xid) == 2) {
    data->device = (nvmlDevice_t)(size_t)(i + 1);
    data->eventType = nvmlEventTypeXidCriticalError;
    data->eventData = xid;
    return NVML_SUCCESS;
  }
  return NVML_ERROR_TIMEOUT;
}
nvmlReturn_t nvmlEventSetWait(nvmlEventSet_t s, nvmlEventData_t *data, unsigned ms) { return wait_impl(data); }
nvmlReturn_t nvmlEventSetWait_v2(nvmlEventSet_t s, nvmlEventData_t *data, unsigned ms) { return wait_impl(data); }
