/*
 * gpushare_b200.h — C ABI of libgpushare_b200.so.
 *
 * This is the inner drop-in boundary of SURVEY.md §8(b): every entry point replaces one or
 * more `nvml.*` call sites of the reference's inventory + health path (all citations are
 * relative to the reference tree):
 *
 *   pkg/gpu/nvidia/gpumanager.go:36,41      nvml.Init / nvml.Shutdown      -> gsb_init / gsb_shutdown
 *   pkg/gpu/nvidia/nvidia.go:47-51,54       nvml.GetDeviceCount            -> gsb_device_count
 *   pkg/gpu/nvidia/nvidia.go:60-71          nvml.NewDevice (UUID/Path/Mem) -> gsb_device_info_get
 *   pkg/gpu/nvidia/nvidia.go:34-45          setGPUMemory / getGPUMemory    -> gsb_slices
 *   pkg/gpu/nvidia/nvidia.go:26-32          generate/extract fake IDs      -> gsb_fake_device_id / gsb_real_device_id
 *   pkg/gpu/nvidia/nvidia.go:100-152        watchXIDs (NewEventSet, RegisterEventForDevice,
 *                                           WaitForEvent, DeleteEventSet)  -> gsb_health_* + gsb_probe*
 *   pkg/gpu/nvidia/server.go:172-185 + vendor/.../v1beta1/api.pb.go:794-843 (gogo MarshalTo)
 *                                                                          -> gsb_encode_list_and_watch
 *   pkg/gpu/nvidia/server.go:150-169        RegisterRequest                -> gsb_encode_register_request
 *   pkg/gpu/nvidia/allocate.go:24-198       Allocate / buildErrResponse    -> gsb_allocate
 *
 * Conventions (cgo / ctypes friendly): plain C, no callbacks, caller-allocated outputs,
 * `int` return (0 = GSB_OK, negative = gsb_status), text via gsb_strerror(). All functions are
 * thread-safe after gsb_init() returned GSB_OK; gsb_probe() blocks the calling thread only.
 * Nothing returned through a pointer outlives the call. There is no CPU fallback: with no
 * usable sm_100 device / driver every device entry point fails with a negative status.
 */
#ifndef GPUSHARE_B200_H_
#define GPUSHARE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_ABI_VERSION 2u /* 2: inventory snapshot + options, transient window, watchdog; gsb_cycle_result grew */

/* = NVML_DEVICE_UUID_BUFFER_SIZE (vendor/.../nvml/nvml.h:1567), used by bindings.go:20 szUUID */
#define GSB_UUID_BUFFER_SIZE 80
#define GSB_BUSID_BUFFER_SIZE 32
#define GSB_MAX_DEVICES 64
/* device-plugin API limit on Device.ID (v1beta1/api.proto:82-85) */
#define GSB_DEVICE_ID_MAX 63

typedef enum gsb_status {
  GSB_OK = 0,
  GSB_ERR_NOT_INITIALIZED = -1,
  GSB_ERR_INVALID_ARGUMENT = -2,
  /* ≙ NVML_ERROR_LIBRARY_NOT_FOUND -> "could not load NVML library" (bindings.go:60-66) */
  GSB_ERR_LIBRARY_NOT_FOUND = -3,
  GSB_ERR_DRIVER = -4,            /* CUDA driver / runtime call failed; see gsb_last_error() */
  GSB_ERR_NVML = -5,              /* NVML call failed; text is "nvml: <nvmlErrorString>" (bindings.go:52-58) */
  GSB_ERR_NO_DEVICE = -6,         /* index out of range / no such UUID ("nvml: device not found", bindings.go:127) */
  GSB_ERR_IDENTITY_MISMATCH = -7, /* CUDA-side and NVML-side identity of one device disagree */
  GSB_ERR_BUFFER_TOO_SMALL = -8,
  GSB_ERR_OUT_OF_MEMORY = -9,
  GSB_ERR_UNSUPPORTED_ARCH = -10, /* device is not compute capability 10.x: no kernel image, no fallback */
  GSB_ERR_TIMEOUT = -11,          /* gsb_health_wait: no event within timeout (≙ NVML_ERROR_TIMEOUT) */
  GSB_ERR_NO_ARENA = -12,         /* gsb_probe before gsb_arena_create */
  GSB_ERR_MALFORMED = -13,        /* protobuf decode failure */
  GSB_ERR_STOPPED = -14           /* health queue closed by gsb_health_stop */
} gsb_status;

/* ---- lifecycle (a1) -------------------------------------------------------------------- */

int gsb_abi_version(void);
/* dlopen libcuda.so.1 + libnvidia-ml.so.1, cuInit, nvmlInit_v2, enumerate devices in PCI-bus
 * (= NVML index) order, cross-check CUDA uuid/bus id against NVML uuid/bus id. Idempotent. */
int gsb_init(void);
int gsb_shutdown(void);
const char *gsb_strerror(int status);
/* thread-local detail of the last failing call on this thread; copies a NUL-terminated string */
int gsb_last_error(char *buf, size_t cap);

/* ---- inventory (a2-a7) ------------------------------------------------------------------ */

typedef struct gsb_device_info {
  char uuid[GSB_UUID_BUFFER_SIZE];     /* "GPU-xxxxxxxx-xxxx-xxxx-xxxx-xxxxxxxxxxxx" = nvml Device.UUID */
  char bus_id[GSB_BUSID_BUFFER_SIZE];  /* NVML busId form "00000000:1B:00.0" */
  uint32_t index;                      /* NVML index == position in PCI-bus order */
  uint32_t minor;                      /* /dev/nvidia<minor>  (nvidia.go:65) */
  int32_t cuda_ordinal;                /* ordinal the CUDA driver gave this device in this process */
  uint32_t sm_count;
  uint32_t cc_major, cc_minor;
  uint32_t reserved0;
  uint64_t total_bytes;                /* nvmlMemory_t.total (v1 struct, nvml.h:178-183): the reference's source */
  uint64_t total_mib;                  /* total_bytes / (1024*1024)   (bindings.go:346-349) */
  uint64_t free_bytes;                 /* nvmlMemory_t.free at query time */
  uint64_t cuda_total_bytes;           /* cuDeviceTotalMem, informational */
} gsb_device_info;

int gsb_device_count(uint32_t *n);
/* ≙ nvml.NewDevice(idx) at plugin (re)start (server.go:39 -> nvidia.go:60): a LIVE query — UUID, minor and
 * nvmlMemory_t re-read from NVML, identity re-verified against the CUDA driver — nothing served from a cache.
 * It also rewrites the device's inventory snapshot (below). */
int gsb_device_info_get(uint32_t idx, gsb_device_info *out);

/*
 * Inventory snapshot. The reference asks NVML for identity and memory ONCE per plugin (re)start (server.go:39,
 * nvidia.go:53-89) and never again while serving; its steady state is a blocked nvmlEventSetWait. Here gsb_init
 * and every gsb_device_info_get / gsb_inventory_refresh store NVML's answer per device, and the steady-state
 * cycle (gsb_cycle / gsb_cycle_all) reads identity + total from that snapshot, re-validating the identity on every
 * cycle against the CUDA driver (cuDeviceGetUuid: user-mode, no driver-wide lock). nvmlDeviceGetMemoryInfo on a
 * 0.3 ms cadence convoys with every other NVML client of the node (nvidia-smi, dcgm-exporter): 6 us idle, up to
 * 2.3 ms contended (BENCH_r01). While health is running a low-rate refresher (GSB_OPT_INVENTORY_REFRESH_MS) re-asks
 * NVML off the cycle's path and raises GSB_EVENT_INVENTORY if the answer changed.
 */
#define GSB_ALL_DEVICES 0xFFFFFFFFu
int gsb_inventory_refresh(uint32_t idx);  /* idx or GSB_ALL_DEVICES: live NVML query -> snapshot */
/* the snapshot as the cycle sees it (identity re-validated), and how old it is */
int gsb_inventory_snapshot(uint32_t idx, gsb_device_info *out, uint64_t *age_ns);

/* ---- options ------------------------------------------------------------------------------ */
enum {
  GSB_OPT_INVENTORY_POLICY = 1,          /* GSB_INVENTORY_SNAPSHOT (default) | GSB_INVENTORY_LIVE */
  GSB_OPT_WAIT_SPIN_US = 2,              /* completion wait: spin budget before sleeping (default 2000; prober
                                            threads with a period >= 10 ms never spin) */
  GSB_OPT_WATCHDOG_MS = 3,               /* completion watchdog of a probe launch (default 2000, + 1 ns per 10
                                            window bytes); 0 = off (waits for ever) */
  GSB_OPT_INVENTORY_REFRESH_MS = 4,      /* low-rate NVML refresh while health runs (default 5000; 0 = never) */
  GSB_OPT_TRANSIENT_KEEP_FREE_BYTES = 5, /* HBM a transient probe window never takes (default 1 GiB) */
  GSB_OPT_SWEEP_EVERY_CYCLES = 6         /* prober, devices without a standing arena: every Nth cycle walks ALL the HBM
                                            that is allocatable at that moment (minus the keep-free margin) instead of
                                            one window, then gives it back — SURVEY §7 hard-part 2's "full walk at idle";
                                            0 = never (default) */
};
enum {
  GSB_INVENTORY_SNAPSHOT = 0, /* cycle: NVML's (re)start-time answer + per-cycle CUDA-side identity check */
  GSB_INVENTORY_LIVE = 1      /* cycle: fresh NVML UUID/minor/MemoryInfo every cycle (round-1 behaviour) */
};
int gsb_set_option(uint32_t key, uint64_t value);
int gsb_get_option(uint32_t key, uint64_t *value);
/* setGPUMemory (nvidia.go:34-41): unit_gib != 0 -> total_mib / 1024, else total_mib. Pure. */
uint32_t gsb_slices(uint64_t total_mib, int unit_gib);
/* "<uuid>-_-<j>" (nvidia.go:26-28). Returns length written (excluding NUL) or negative status. */
int gsb_fake_device_id(const char *uuid, uint32_t j, char *buf, size_t cap);
/* strings.Split(id, "-_-")[0] (nvidia.go:30-32). Returns length or negative status. */
int gsb_real_device_id(const char *fake_id, char *buf, size_t cap);

/* ---- wire encoders (a7, a9; gogo-identical bytes) ---------------------------------------- */

/*
 * ListAndWatchResponse{devices: for g in [0,n_gpus), j in [0,slices): Device{ID:"<uuid[g]>-_-<j>",
 * health: unhealthy_bits bit (g*slices+j) ? "Unhealthy" : "Healthy"}} — the list getDevices()
 * builds (nvidia.go:73-85) marshalled like api.pb.go:794-843. `unhealthy_bits` may be NULL (all
 * healthy); otherwise it holds ceil(n_gpus*slices/8) bytes, LSB-first. Returns bytes written, or
 * with buf==NULL the size required, or a negative status.
 */
int64_t gsb_encode_list_and_watch(const char *const *uuids, uint32_t n_gpus, uint32_t slices,
                                  const uint8_t *unhealthy_bits, uint8_t *buf, size_t cap);
/* RegisterRequest{version, endpoint, resource_name} (server.go:158-162; api.pb.go:730-792). */
int64_t gsb_encode_register_request(const char *version, const char *endpoint,
                                    const char *resource_name, uint8_t *buf, size_t cap);

/* ---- HBM probe (replaces the passive XID wait with an active walk) --------- */

enum {
  GSB_OP_FILL = 1,          /* write pattern(seed_write); traffic = W */
  GSB_OP_VERIFY = 2,        /* read, compare with pattern(seed_expect); traffic = W */
  GSB_OP_VERIFY_REFILL = 3  /* read+compare(seed_expect), write pattern(seed_write); traffic = 2W */
};
enum {
  GSB_VARIANT_AUTO = 0,
  GSB_VARIANT_DIRECT = 1,   /* ld.global.v4 -> registers -> st.global.v4 (control: no staging) */
  GSB_VARIANT_CPASYNC = 2,  /* cp.async 16 B -> shared ring -> ld.shared.v4 -> st.global.v4 */
  GSB_VARIANT_BULK = 3,     /* cp.async.bulk (TMA 1-D) -> shared ring -> ld.shared.v4 / st.shared.v4
                               -> cp.async.bulk shared->global; one ring per CTA, CTA barrier per tile */
  GSB_VARIANT_BULKW = 4,    /* same data path, one private ring + mbarriers per WARP: no CTA barrier */
  GSB_VARIANT_BULKD = 5     /* BULK with a dynamic (atomic-counter) tile scheduler instead of the static stride */
};
enum {
  GSB_PROBE_TIMED = 1u,       /* bracket the launch with CUDA events, fill kernel_ns */
  GSB_PROBE_SEED_TABLE = 2u   /* expected seed per granule from the arena's generation table */
};

typedef struct gsb_probe_cfg {
  uint32_t op;            /* GSB_OP_* */
  uint32_t variant;       /* GSB_VARIANT_* */
  uint64_t window_offset; /* bytes from arena start, multiple of 16 */
  uint64_t window_bytes;  /* multiple of 16; 0 = to the end of the arena */
  uint32_t seed_expect;
  uint32_t seed_write;
  uint32_t grid_ctas;     /* 0 = auto (resident CTAs per SM x SM count) */
  uint32_t flags;         /* GSB_PROBE_* */
} gsb_probe_cfg;

typedef struct gsb_probe_result {
  int32_t status;            /* gsb_status of this device's probe */
  uint32_t variant;          /* variant actually launched */
  uint64_t bytes_walked;     /* window size actually walked */
  uint64_t bytes_read;       /* algorithmic bytes loaded  (0 for FILL) */
  uint64_t bytes_written;    /* algorithmic bytes stored  (0 for VERIFY) */
  uint64_t mismatch_words;   /* 16-byte words with >= 1 differing bit */
  uint64_t mismatch_bits;    /* total differing bits */
  uint64_t first_bad_offset; /* lowest arena byte offset of a mismatching word; UINT64_MAX if none */
  uint32_t checksum_xor;     /* XOR of every 32-bit lane observed (read ops) or written (FILL) */
  uint32_t checksum_sum;     /* wrapping 32-bit sum of the same lanes */
  uint64_t kernel_ns;        /* CUDA-event time of the launch (GSB_PROBE_TIMED), else 0 */
  uint64_t wall_ns;          /* host monotonic time of the whole call */
  uint32_t grid_ctas;
  uint32_t block_threads;
} gsb_probe_result;

/*
 * Arena = the device memory the probe walks: every byte the driver will actually hand out right
 * now (max_bytes == 0), found by mapping physical chunks (cuMemCreate, 2 MiB granularity) into one
 * reserved VA range until the driver refuses; `keep_free_bytes` is left unallocated for tenants.
 * *arena_bytes is the "actually allocatable" figure.
 */
int gsb_arena_create(uint32_t idx, uint64_t max_bytes, uint64_t keep_free_bytes, uint64_t *arena_bytes);
int gsb_arena_destroy(uint32_t idx);
int gsb_arena_bytes(uint32_t idx, uint64_t *arena_bytes);
int gsb_probe(uint32_t idx, const gsb_probe_cfg *cfg, gsb_probe_result *out);
/* Same cfg on n devices concurrently: one host thread, one primary context and one non-blocking
 * stream per device. Returns GSB_OK iff every results[i].status == GSB_OK. */
int gsb_probe_all(uint32_t n, const uint32_t *idxs, const gsb_probe_cfg *cfg, gsb_probe_result *results);
/* test hooks: raw access to arena bytes (fault injection / read-back by the parity tests) */
int gsb_arena_read(uint32_t idx, uint64_t offset, void *dst, uint64_t bytes);
int gsb_arena_write(uint32_t idx, uint64_t offset, const void *src, uint64_t bytes);
/* test hook: enqueue a kernel that keeps the device's probe stream busy for `ms` milliseconds — a stand-in for a
 * wedged GPU, to exercise the completion watchdog */
int gsb_test_stall(uint32_t idx, uint32_t ms);
/* test hook: overwrite the snapshot's total (as if NVML had answered differently at the last (re)start), to exercise the
 * off-path refresher: its next pass must raise GSB_EVENT_INVENTORY / GSB_INVENTORY_TOTAL_CHANGED for this device */
int gsb_test_skew_snapshot(uint32_t idx, uint64_t total_bytes);

/*
 * One inventory + health-probe cycle of one device (the unit of BASELINE.json's metric):
 * identity + total (GSB_OPT_INVENTORY_POLICY: snapshot re-validated on the CUDA side, or a live NVML query)
 * -> slices -> S fake devices -> ListAndWatchResponse bytes into lw_buf -> HBM walk -> verdict.
 *   standing arena (gsb_arena_create was called): VERIFY_REFILL of the arena window that holds slice
 *     (cycle_no mod arena_slices) (window_bytes == 0: the whole arena); traffic 2 * W.
 *   no arena, window_bytes > 0: TRANSIENT window (SURVEY.md §7 hard-part 2) — allocate up to window_bytes
 *     (never the last GSB_OPT_TRANSIENT_KEEP_FREE_BYTES), FILL, VERIFY, free; traffic 2 * W; nothing is held
 *     between cycles. If nothing can be allocated the cycle is inventory-only: rc GSB_OK, healthy unchanged,
 *     probe.status == GSB_ERR_OUT_OF_MEMORY, probe.bytes_walked == 0.
 * A launch that outlives the watchdog returns GSB_ERR_TIMEOUT with healthy == 0.
 */
typedef struct gsb_cycle_result {
  gsb_device_info info;
  uint32_t slices;
  uint32_t healthy;      /* 1 iff this cycle's walk was clean AND no earlier cycle faulted (Unhealthy is sticky,
                            server.go:180) */
  int64_t lw_len;        /* bytes written to lw_buf */
  uint64_t inventory_ns; /* host time of the identity/memory step + encode */
  gsb_probe_result probe;
  uint64_t snapshot_age_ns; /* policy SNAPSHOT: age of NVML's answer the cycle served */
  uint32_t inventory_live;  /* 1 = this cycle queried NVML itself */
  uint32_t transient;       /* 1 = transient window (allocate -> walk -> free) */
} gsb_cycle_result;
int gsb_cycle(uint32_t idx, uint64_t cycle_no, uint64_t window_bytes, int unit_gib, uint32_t variant,
              uint8_t *lw_buf, size_t lw_cap, gsb_cycle_result *out);

/*
 * One NODE cycle: gsb_cycle on every listed device concurrently — one persistent host thread, one
 * primary context and one non-blocking stream per device, no collective — then
 * the single host-side join: the per-device lists concatenated in index order into ONE
 * ListAndWatchResponse (what getDevices builds sequentially, nvidia.go:59-86). The slice count of
 * device idxs[0] is applied to every device, as the reference's process-global gpuMemory does
 * (nvidia.go:70-72). results[i] is device idxs[i]'s cycle; returns bytes written to lw_buf or a
 * negative status.
 */
int64_t gsb_cycle_all(uint32_t n, const uint32_t *idxs, uint64_t cycle_no, uint64_t window_bytes, int unit_gib,
                      uint32_t variant, uint8_t *lw_buf, size_t lw_cap, gsb_cycle_result *results);

/* ---- health events (a10, a11) ------------------------------------------------------------ */

enum {
  GSB_EVENT_XID = 8,        /* = nvmlEventTypeXidCriticalError (nvml.h:1082); edata = XID */
  GSB_EVENT_PROBE = 0x100,  /* active probe verdict; edata = GSB_PROBE_FAULT_* */
  GSB_EVENT_INVENTORY = 0x200 /* the low-rate NVML refresh disagrees with the snapshot; edata = GSB_INVENTORY_*_CHANGED */
};
enum { GSB_PROBE_FAULT_MISMATCH = 1, GSB_PROBE_FAULT_LAUNCH = 2, GSB_PROBE_RECOVERED = 3,
       GSB_PROBE_FAULT_WEDGED = 4 /* the launch outlived the completion watchdog */ };
enum { GSB_INVENTORY_IDENTITY_CHANGED = 1, GSB_INVENTORY_TOTAL_CHANGED = 2 };

typedef struct gsb_event {
  char uuid[GSB_UUID_BUFFER_SIZE]; /* empty => applies to all devices (nvidia.go:138-144) */
  uint64_t etype;
  uint64_t edata;
} gsb_event;

/* Start: one NVML event set with XidCriticalError registered once per *GPU* (the reference does
 * it once per fake device, nvidia.go:104-117 — same resulting set) plus, if probe_period_ms > 0, a
 * prober thread per device running gsb_cycle (standing arena if one exists, else transient windows).
 * A registration that fails with anything but NOT_SUPPORTED fails the call with GSB_ERR_NVML and starts
 * nothing (the reference: log.Fatalf, nvidia.go:114-116); NOT_SUPPORTED queues an Unhealthy event for
 * that GPU (nvidia.go:107-112). */
int gsb_health_start(uint32_t probe_period_ms, uint64_t window_bytes);
int gsb_health_stop(void);
/* ≙ nvml.WaitForEvent(set, timeout) (bindings.go:134-146): GSB_OK + event, or GSB_ERR_TIMEOUT. */
int gsb_health_wait(uint32_t timeout_ms, gsb_event *ev);
/* What the prober thread of one device has done since gsb_health_start (counters restart with it). */
typedef struct gsb_health_stats {
  uint64_t cycles;            /* probe cycles run */
  uint64_t sweeps;            /* of which: whole-free-HBM sweeps (GSB_OPT_SWEEP_EVERY_CYCLES) */
  uint64_t skipped;           /* cycles that found nothing allocatable to probe */
  uint64_t faults;            /* cycles whose walk was not clean (mismatch, failed or wedged launch) */
  uint64_t last_bytes_walked; /* of the most recent cycle */
  uint64_t last_kernel_ns;
  uint64_t last_sweep_bytes;  /* bytes the most recent sweep walked = what was actually allocatable then */
  uint64_t last_sweep_ns;     /* host time of that sweep, allocation and release included */
} gsb_health_stats;
int gsb_health_stats_get(uint32_t idx, gsb_health_stats *out);
/* SURVEY.md §8(f) rank 1, optional: after `clean_cycles` consecutive clean probe cycles a GPU that a PROBE
 * verdict had marked unhealthy is reported again with edata = GSB_PROBE_RECOVERED. 0 (default) keeps the
 * reference's behaviour: Unhealthy is sticky (server.go:180 FIXME). XID faults never recover. */
int gsb_health_set_recovery(uint32_t clean_cycles);
/* test hook: enqueue an event as if the driver had delivered it */
int gsb_health_inject(const gsb_event *ev);
/* nvidia.go:134: XIDs 31, 43, 45 are application errors and do not mark the GPU unhealthy. Pure. */
int gsb_xid_is_benign(uint64_t xid);

/* ---- Allocate (a12, a13): wire bytes in, wire bytes out ---------------------------------- */

typedef struct gsb_pod {
  const char *name;
  const char *ns;
  const char *uid;
  uint64_t gpu_mem_limit;   /* Σ spec.containers[].resources.limits["aliyun.com/gpu-mem"] (podutils.go:122-131) */
  uint64_t assume_time;     /* ALIYUN_COM_GPU_MEM_ASSUME_TIME parsed as uint64, 0 on failure (podutils.go:64-75) */
  int32_t gpu_idx;          /* ALIYUN_COM_GPU_MEM_IDX via Atoi, -1 if absent/unparsable (podutils.go:37-61) */
  uint8_t has_assume_time;  /* annotation key present */
  uint8_t has_assigned;     /* ALIYUN_COM_GPU_MEM_ASSIGNED present */
  uint8_t assigned_is_false;/* value == "false" */
  uint8_t on_node;          /* spec.nodeName == $NODE_NAME (podmanager.go:187-193); others are skipped */
} gsb_pod;

typedef struct gsb_allocate_ctx {
  const char *const *uuids; /* per GPU, NVML order */
  const uint32_t *minors;   /* per GPU: devNameMap[uuid] */
  uint32_t n_gpus;
  uint32_t slices;          /* getGPUMemory() */
  int32_t unit_gib;         /* metric == GiBPrefix */
  int32_t disable_cgpu_isolation;
  int32_t pods_unique;      /* caller guarantees that no two on-node pods share a uid (a table keyed by uid):
                               getPendingPodsInNode's dedupe (podmanager.go:162-212) is then a no-op and is skipped.
                               0 = dedupe as the reference does */
} gsb_allocate_ctx;

enum {
  GSB_ALLOC_MATCHED = 1,     /* pod found; caller must PATCH pod `*pod_index` (allocate.go:130-149) */
  GSB_ALLOC_SINGLE_GPU = 2,  /* no pod matched, one GPU: shortcut response (allocate.go:151-177) */
  GSB_ALLOC_ERR_RESPONSE = 3 /* buildErrResponse (allocate.go:24-39) */
};

/*
 * Decode AllocateRequest wire bytes, run the reference's selection over `pods` (candidate filter
 * podutils.go:78-119, order podmanager.go:241-262, first pod whose limit == request
 * allocate.go:78-88), and encode the AllocateResponse wire bytes (env keys emitted in sorted key
 * order — gogo iterates a Go map, so any order is conformant). Returns
 * GSB_ALLOC_* (>0) or a negative status. `*pod_index` = index into pods of the matched pod or -1.
 */
int gsb_allocate(const gsb_allocate_ctx *ctx, const gsb_pod *pods, uint32_t n_pods,
                 const uint8_t *req, size_t req_len, uint8_t *resp, size_t resp_cap,
                 size_t *resp_len, int32_t *pod_index, uint32_t *pod_req_gpu);
/* Re-encode the error response for a request (PATCH failed after a match: allocate.go:133-148). */
int gsb_allocate_err_response(const gsb_allocate_ctx *ctx, const uint8_t *req, size_t req_len,
                              uint8_t *resp, size_t resp_cap, size_t *resp_len);
/* {"metadata":{"annotations":{"ALIYUN_COM_GPU_MEM_ASSIGNED":"true","ALIYUN_COM_GPU_MEM_ASSUME_TIME":"<ns>"}}}
 * (podutils.go:27-35). Returns length or negative status. */
int gsb_patch_assigned_body(uint64_t now_unix_ns, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* GPUSHARE_B200_H_ */
