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
	gsbEventXID   = uint64(C.GSB_EVENT_XID)   // == nvml.XidCriticalError
	gsbEventProbe = uint64(C.GSB_EVENT_PROBE) // verdict of the active HBM probe
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

// gsbCycle runs one inventory + health-probe cycle of one device; lw receives the ListAndWatchResponse bytes.
func gsbCycle(idx uint, cycle uint64, windowBytes uint64, lw []byte) (healthy bool, lwLen int, err error) {
	var res C.gsb_cycle_result
	rc := C.gsb_cycle(C.uint32_t(idx), C.uint64_t(cycle), C.uint64_t(windowBytes), 1, C.GSB_VARIANT_AUTO,
		(*C.uint8_t)(unsafe.Pointer(&lw[0])), C.size_t(len(lw)), &res)
	return res.healthy == 1, int(res.lw_len), gsbErr(rc)
}
