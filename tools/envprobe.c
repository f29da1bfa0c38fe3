/*
 * envprobe — first-contact facts about the GPU box.
 *
 * Prints, per device: NVML index, PCI bus id, UUID, minor, NVML memory totals (v1 and v2
 * structs), and for the CUDA driver API: ordinal, bus id, UUID, cuDeviceTotalMem and
 * cuMemGetInfo inside a primary context. Not part of the product; diagnostics only.
 *
 * Build: gcc -O2 -o envprobe envprobe.c -I/usr/local/cuda/include -ldl
 */
#include <cuda.h>
#include <dlfcn.h>
#include <nvml.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#define SYM(lib, name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name)

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(void) {
  void *ml = dlopen("libnvidia-ml.so.1", RTLD_LAZY | RTLD_GLOBAL);
  void *cu = dlopen("libcuda.so.1", RTLD_LAZY | RTLD_GLOBAL);
  printf("dlopen nvml=%p cuda=%p\n", ml, cu);
  if (!ml || !cu) return 1;

  nvmlReturn_t (*p_init)(void) = (nvmlReturn_t(*)(void))dlsym(ml, "nvmlInit_v2");
  nvmlReturn_t (*p_count)(unsigned *) = (nvmlReturn_t(*)(unsigned *))dlsym(ml, "nvmlDeviceGetCount_v2");
  nvmlReturn_t (*p_hbi)(unsigned, nvmlDevice_t *) =
      (nvmlReturn_t(*)(unsigned, nvmlDevice_t *))dlsym(ml, "nvmlDeviceGetHandleByIndex_v2");
  nvmlReturn_t (*p_uuid)(nvmlDevice_t, char *, unsigned) =
      (nvmlReturn_t(*)(nvmlDevice_t, char *, unsigned))dlsym(ml, "nvmlDeviceGetUUID");
  nvmlReturn_t (*p_minor)(nvmlDevice_t, unsigned *) =
      (nvmlReturn_t(*)(nvmlDevice_t, unsigned *))dlsym(ml, "nvmlDeviceGetMinorNumber");
  nvmlReturn_t (*p_mem1)(nvmlDevice_t, nvmlMemory_t *) =
      (nvmlReturn_t(*)(nvmlDevice_t, nvmlMemory_t *))dlsym(ml, "nvmlDeviceGetMemoryInfo");
  nvmlReturn_t (*p_mem2)(nvmlDevice_t, nvmlMemory_v2_t *) =
      (nvmlReturn_t(*)(nvmlDevice_t, nvmlMemory_v2_t *))dlsym(ml, "nvmlDeviceGetMemoryInfo_v2");
  nvmlReturn_t (*p_pci)(nvmlDevice_t, nvmlPciInfo_t *) =
      (nvmlReturn_t(*)(nvmlDevice_t, nvmlPciInfo_t *))dlsym(ml, "nvmlDeviceGetPciInfo_v3");
  nvmlReturn_t (*p_drv)(char *, unsigned) =
      (nvmlReturn_t(*)(char *, unsigned))dlsym(ml, "nvmlSystemGetDriverVersion");

  double t0 = now_us();
  nvmlReturn_t r = p_init();
  printf("nvmlInit_v2 -> %d  (%.0f us)\n", r, now_us() - t0);
  char drv[96] = {0};
  p_drv(drv, sizeof drv);
  printf("driver %s\n", drv);
  unsigned n = 0;
  p_count(&n);
  printf("nvml device count %u\n", n);
  for (unsigned i = 0; i < n; i++) {
    nvmlDevice_t d;
    p_hbi(i, &d);
    char uuid[96] = {0};
    unsigned minor = 9999;
    nvmlMemory_t m1;
    nvmlMemory_v2_t m2;
    memset(&m1, 0, sizeof m1);
    memset(&m2, 0, sizeof m2);
    m2.version = nvmlMemory_v2;
    nvmlPciInfo_t pci;
    memset(&pci, 0, sizeof pci);
    p_uuid(d, uuid, sizeof uuid);
    p_minor(d, &minor);
    double a = now_us();
    nvmlReturn_t r1 = p_mem1(d, &m1);
    double b = now_us();
    nvmlReturn_t r2 = p_mem2 ? p_mem2(d, &m2) : 999;
    p_pci(d, &pci);
    printf("nvml[%u] bus=%s uuid=%s minor=%u\n", i, pci.busId, uuid, minor);
    printf("  v1(r=%d, %.1f us): total=%llu free=%llu used=%llu  MiB=%llu GiB=%llu\n", r1, b - a,
           m1.total, m1.free, m1.used, m1.total / 1048576ULL, m1.total / 1048576ULL / 1024ULL);
    printf("  v2(r=%d): total=%llu reserved=%llu free=%llu used=%llu MiB=%llu\n", r2, m2.total,
           m2.reserved, m2.free, m2.used, m2.total / 1048576ULL);
  }

  CUresult (*c_init)(unsigned) = (CUresult(*)(unsigned))dlsym(cu, "cuInit");
  CUresult (*c_count)(int *) = (CUresult(*)(int *))dlsym(cu, "cuDeviceGetCount");
  CUresult (*c_get)(CUdevice *, int) = (CUresult(*)(CUdevice *, int))dlsym(cu, "cuDeviceGet");
  CUresult (*c_uuid)(CUuuid *, CUdevice) = (CUresult(*)(CUuuid *, CUdevice))dlsym(cu, "cuDeviceGetUuid_v2");
  if (!c_uuid) c_uuid = (CUresult(*)(CUuuid *, CUdevice))dlsym(cu, "cuDeviceGetUuid");
  CUresult (*c_bus)(char *, int, CUdevice) = (CUresult(*)(char *, int, CUdevice))dlsym(cu, "cuDeviceGetPCIBusId");
  CUresult (*c_tot)(size_t *, CUdevice) = (CUresult(*)(size_t *, CUdevice))dlsym(cu, "cuDeviceTotalMem_v2");
  CUresult (*c_retain)(CUcontext *, CUdevice) = (CUresult(*)(CUcontext *, CUdevice))dlsym(cu, "cuDevicePrimaryCtxRetain");
  CUresult (*c_setcur)(CUcontext) = (CUresult(*)(CUcontext))dlsym(cu, "cuCtxSetCurrent");
  CUresult (*c_meminfo)(size_t *, size_t *) = (CUresult(*)(size_t *, size_t *))dlsym(cu, "cuMemGetInfo_v2");
  CUresult (*c_attr)(int *, CUdevice_attribute, CUdevice) =
      (CUresult(*)(int *, CUdevice_attribute, CUdevice))dlsym(cu, "cuDeviceGetAttribute");

  t0 = now_us();
  CUresult cr = c_init(0);
  printf("cuInit -> %d (%.0f us)\n", cr, now_us() - t0);
  int cn = 0;
  c_count(&cn);
  printf("cuda device count %d\n", cn);
  for (int i = 0; i < cn; i++) {
    CUdevice dev;
    c_get(&dev, i);
    CUuuid u;
    c_uuid(&u, dev);
    char bus[32] = {0};
    c_bus(bus, sizeof bus, dev);
    size_t tot = 0;
    c_tot(&tot, dev);
    int sms = 0, vmm = 0, l2 = 0;
    c_attr(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev);
    c_attr(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
    c_attr(&l2, CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE, dev);
    printf("cuda[%d] bus=%s uuid=", i, bus);
    for (int k = 0; k < 16; k++) printf("%02x", (unsigned char)u.bytes[k]);
    printf(" totalMem=%zu (MiB=%zu GiB=%zu) sms=%d vmm=%d l2=%d\n", tot, tot >> 20, tot >> 30, sms, vmm, l2);
    CUcontext ctx;
    t0 = now_us();
    cr = c_retain(&ctx, dev);
    c_setcur(ctx);
    size_t fr = 0, tt = 0;
    c_meminfo(&fr, &tt);
    printf("  ctx retain r=%d (%.0f us) cuMemGetInfo free=%zu total=%zu (free MiB=%zu)\n", cr,
           now_us() - t0, fr, tt, fr >> 20);
  }
  /* NVML view again, now that contexts exist */
  for (unsigned i = 0; i < n; i++) {
    nvmlDevice_t d;
    p_hbi(i, &d);
    nvmlMemory_t m1;
    p_mem1(d, &m1);
    printf("nvml[%u] after ctx: total=%llu free=%llu used=%llu\n", i, m1.total, m1.free, m1.used);
  }
  return 0;
}
