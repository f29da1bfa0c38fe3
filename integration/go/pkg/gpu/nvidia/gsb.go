// +build linux

package nvidia

// Thin cgo shim over libgpushare_b200.so (include/gpushare_b200.h). NOT compiled in this repository:
// there is no Go toolchain in the build image. The same entry points are exercised by the ctypes
// binding and by the native daemon.

// #cgo CFLAGS: -I${SRCDIR}/../../../include
// #cgo LDFLAGS: -L${SRCDIR}/../../../lib -lgpushare_b200 -Wl,-rpath,$ORIGIN/../lib
// #include <stdlib.h>
// #include "gpushare_b200.h"
import "C"

import (
	"errors"
	"fmt"
	"unsafe"
)

func gsbErr(rc C.int) error {
	if rc >= 0 {
		return nil
	}
	var buf [512]C.char
	C.gsb_last_error(&buf[0], C.size_t(len(buf)))
	if rc == C.GSB_ERR_LIBRARY_NOT_FOUND {
		// same text nvml.Init() returns (vendor/.../nvml/bindings.go:63)
		return errors.New("could not load NVML library")
	}
	return fmt.Errorf("%s: %s", C.GoString(C.gsb_strerror(rc)), C.GoString(&buf[0]))
}

// gsbInit / gsbShutdown replace nvml.Init / nvml.Shutdown (gpumanager.go:36,41).
func gsbInit() error     { return gsbErr(C.gsb_init()) }
func gsbShutdown() error { return gsbErr(C.gsb_shutdown()) }

func gsbDeviceCount() (uint, error) {
	var n C.uint32_t
	err := gsbErr(C.gsb_device_count(&n))
	return uint(n), err
}

// gsbDevice carries the three fields of nvml.Device the plugin consumes (nvidia.go:64-71).
type gsbDevice struct {
	UUID   string
	Path   string // "/dev/nvidia<minor>"
	Memory uint64 // MiB == *nvml.Device.Memory
}

func gsbNewDevice(idx uint) (*gsbDevice, error) {
	var info C.gsb_device_info
	if err := gsbErr(C.gsb_device_info_get(C.uint32_t(idx), &info)); err != nil {
		return nil, err
	}
	return &gsbDevice{
		UUID:   C.GoString(&info.uuid[0]),
		Path:   fmt.Sprintf("/dev/nvidia%d", uint(info.minor)),
		Memory: uint64(info.total_mib),
	}, nil
}

// gsbRefreshInventory re-reads every GPU's identity + memory from NVML into the library's snapshot. gsbNewDevice does
// that per device already, so a plugin that builds its list through gsbNewDevice at every (re)start — as
// NewNvidiaDevicePlugin does (server.go:39) — needs it only to force a refresh outside a restart.
func gsbRefreshInventory() error { return gsbErr(C.gsb_inventory_refresh(C.GSB_ALL_DEVICES)) }

// Options (include/gpushare_b200.h, GSB_OPT_*): inventory policy of the cycle, completion-wait spin budget, probe
// watchdog, off-path NVML refresh period, HBM a transient probe window never takes.
func gsbSetOption(key uint32, value uint64) error {
	return gsbErr(C.gsb_set_option(C.uint32_t(key), C.uint64_t(value)))
}

// gsbSlices is setGPUMemory's arithmetic (nvidia.go:34-41).
func gsbSlices(totalMiB uint64, gib bool) uint {
	unit := C.int(0)
	if gib {
		unit = 1
	}
	return uint(C.gsb_slices(C.uint64_t(totalMiB), unit))
}

type gsbEvent struct {
	UUID  string
	Etype uint64
	Edata uint64
}

const (
	gsbEventXID       = uint64(C.GSB_EVENT_XID)       // == nvml.XidCriticalError
	gsbEventProbe     = uint64(C.GSB_EVENT_PROBE)     // verdict of the active HBM probe (mismatch, failed launch, wedged)
	gsbEventInventory = uint64(C.GSB_EVENT_INVENTORY) // the off-path NVML refresh disagrees with what is advertised
)

// gsbHealthStart replaces NewEventSet + the RegisterEventForDevice loop (nvidia.go:101-117).
func gsbHealthStart(probePeriodMs uint, windowBytes uint64) error {
	return gsbErr(C.gsb_health_start(C.uint32_t(probePeriodMs), C.uint64_t(windowBytes)))
}

func gsbHealthStop() { C.gsb_health_stop() }

// gsbWaitForEvent replaces nvml.WaitForEvent (nvidia.go:126). (nil, nil) == timeout.
// cgo releases the P while the call blocks: one OS thread, like the reference's.
func gsbWaitForEvent(timeoutMs uint) (*gsbEvent, error) {
	var ev C.gsb_event
	rc := C.gsb_health_wait(C.uint32_t(timeoutMs), &ev)
	if rc == C.GSB_ERR_TIMEOUT || rc == C.GSB_ERR_STOPPED {
		return nil, nil
	}
	if err := gsbErr(rc); err != nil {
		return nil, err
	}
	return &gsbEvent{C.GoString(&ev.uuid[0]), uint64(ev.etype), uint64(ev.edata)}, nil
}

func gsbXidIsBenign(xid uint64) bool { return C.gsb_xid_is_benign(C.uint64_t(xid)) != 0 }

// gsbArenaCreate maps the memory the probe walks; returns what was actually allocatable.
func gsbArenaCreate(idx uint, maxBytes, keepFreeBytes uint64) (uint64, error) {
	var got C.uint64_t
	err := gsbErr(C.gsb_arena_create(C.uint32_t(idx), C.uint64_t(maxBytes), C.uint64_t(keepFreeBytes), &got))
	return uint64(got), err
}

// gsbCycle runs one inventory + health-probe cycle of one device; lw receives the ListAndWatchResponse bytes. With no
// arena on the device and windowBytes > 0 the cycle probes a transient window (allocate -> fill -> verify -> free).
func gsbCycle(idx uint, cycle uint64, windowBytes uint64, lw []byte) (healthy bool, lwLen int, err error) {
	var res C.gsb_cycle_result
	rc := C.gsb_cycle(C.uint32_t(idx), C.uint64_t(cycle), C.uint64_t(windowBytes), 1, C.GSB_VARIANT_AUTO,
		(*C.uint8_t)(unsafe.Pointer(&lw[0])), C.size_t(len(lw)), &res)
	return res.healthy == 1, int(res.lw_len), gsbErr(rc)
}

// gsbPod is the handful of fields Allocate's pod selection reads (podutils.go:37-131, podmanager.go:187-201).
type gsbPod struct {
	Name, Namespace, UID string
	GPUMemLimit          uint64 // sum of spec.containers[].resources.limits["aliyun.com/gpu-mem"]
	AssumeTime           uint64 // ALIYUN_COM_GPU_MEM_ASSUME_TIME, 0 when absent or unparsable
	GPUIdx               int32  // ALIYUN_COM_GPU_MEM_IDX, -1 when absent or unparsable
	HasAssumeTime        bool
	HasAssigned          bool
	AssignedIsFalse      bool
	OnNode               bool
}

const (
	gsbAllocMatched     = int(C.GSB_ALLOC_MATCHED)      // PATCH pods[podIndex], then answer with resp
	gsbAllocSingleGPU   = int(C.GSB_ALLOC_SINGLE_GPU)   // allocate.go:151-177
	gsbAllocErrResponse = int(C.GSB_ALLOC_ERR_RESPONSE) // allocate.go:24-39
)

// gsbAllocate is the optional fast path for allocate.go:54-128: request bytes in (the gogo-marshalled
// AllocateRequest), response bytes out, and which pending pod to PATCH. The LIST before it and the PATCH after it
// stay in Go (podmanager.go:125-160, allocate.go:130-149). uuids/minors are devNameMap in NVML order.
func gsbAllocate(uuids []string, minors []uint32, slices uint, gib, disableCGPU bool, pods []gsbPod, req []byte) (kind int, resp []byte, podIndex int, err error) {
	cu := make([]*C.char, len(uuids))
	for i, u := range uuids {
		cu[i] = C.CString(u)
		defer C.free(unsafe.Pointer(cu[i]))
	}
	cm := make([]C.uint32_t, len(minors))
	for i, m := range minors {
		cm[i] = C.uint32_t(m)
	}
	b2i := func(b bool) C.int32_t {
		if b {
			return 1
		}
		return 0
	}
	// C memory for everything the C side keeps pointers to during the call (cgo forbids Go pointers to Go pointers)
	ctx := (*C.gsb_allocate_ctx)(C.calloc(1, C.size_t(unsafe.Sizeof(C.gsb_allocate_ctx{}))))
	defer C.free(unsafe.Pointer(ctx))
	cuArr := (**C.char)(C.calloc(C.size_t(len(cu)+1), C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(cuArr))
	cmArr := (*C.uint32_t)(C.calloc(C.size_t(len(cm)+1), 4))
	defer C.free(unsafe.Pointer(cmArr))
	copy(unsafe.Slice(cuArr, len(cu)), cu)
	copy(unsafe.Slice(cmArr, len(cm)), cm)
	ctx.uuids, ctx.minors, ctx.n_gpus = cuArr, cmArr, C.uint32_t(len(uuids))
	ctx.slices, ctx.unit_gib, ctx.disable_cgpu_isolation = C.uint32_t(slices), b2i(gib), b2i(disableCGPU)

	cp := (*C.gsb_pod)(C.calloc(C.size_t(len(pods)+1), C.size_t(unsafe.Sizeof(C.gsb_pod{}))))
	defer C.free(unsafe.Pointer(cp))
	cps := unsafe.Slice(cp, len(pods))
	b2u := func(b bool) C.uint8_t {
		if b {
			return 1
		}
		return 0
	}
	for i, p := range pods {
		name, ns, uid := C.CString(p.Name), C.CString(p.Namespace), C.CString(p.UID)
		defer C.free(unsafe.Pointer(name))
		defer C.free(unsafe.Pointer(ns))
		defer C.free(unsafe.Pointer(uid))
		cps[i].name, cps[i].ns, cps[i].uid = name, ns, uid
		cps[i].gpu_mem_limit, cps[i].assume_time, cps[i].gpu_idx = C.uint64_t(p.GPUMemLimit), C.uint64_t(p.AssumeTime), C.int32_t(p.GPUIdx)
		cps[i].has_assume_time, cps[i].has_assigned = b2u(p.HasAssumeTime), b2u(p.HasAssigned)
		cps[i].assigned_is_false, cps[i].on_node = b2u(p.AssignedIsFalse), b2u(p.OnNode)
	}
	out := make([]byte, 64<<10)
	var n C.size_t
	var idx C.int32_t
	var podReq C.uint32_t
	var reqPtr *C.uint8_t
	if len(req) > 0 {
		reqPtr = (*C.uint8_t)(unsafe.Pointer(&req[0]))
	}
	rc := C.gsb_allocate(ctx, cp, C.uint32_t(len(pods)), reqPtr, C.size_t(len(req)),
		(*C.uint8_t)(unsafe.Pointer(&out[0])), C.size_t(len(out)), &n, &idx, &podReq)
	if rc < 0 {
		return 0, nil, -1, gsbErr(rc)
	}
	return int(rc), out[:int(n)], int(idx), nil
}

// gsbEncodeListAndWatch produces the bytes gogo's Marshal would for the list getDevices() builds (server.go:173,182);
// unhealthy is a bitmap over (gpu*slices + j). Use with a raw-bytes grpc codec, or keep gogo: the bytes are identical.
func gsbEncodeListAndWatch(uuids []string, slices uint, unhealthy []byte) ([]byte, error) {
	cu := (**C.char)(C.calloc(C.size_t(len(uuids)+1), C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(cu))
	for i, u := range uuids {
		s := C.CString(u)
		defer C.free(unsafe.Pointer(s))
		unsafe.Slice(cu, len(uuids))[i] = s
	}
	var bits *C.uint8_t
	if len(unhealthy) > 0 {
		bits = (*C.uint8_t)(unsafe.Pointer(&unhealthy[0]))
	}
	need := C.gsb_encode_list_and_watch(cu, C.uint32_t(len(uuids)), C.uint32_t(slices), bits, nil, 0)
	if need < 0 && C.int(need) != C.GSB_ERR_BUFFER_TOO_SMALL {
		return nil, gsbErr(C.int(need))
	}
	buf := make([]byte, (len(uuids)*int(slices))*64+16)
	n := C.gsb_encode_list_and_watch(cu, C.uint32_t(len(uuids)), C.uint32_t(slices), bits, (*C.uint8_t)(unsafe.Pointer(&buf[0])), C.size_t(len(buf)))
	if n < 0 {
		return nil, gsbErr(C.int(n))
	}
	return buf[:int(n)], nil
}
