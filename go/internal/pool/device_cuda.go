//go:build cuda && cgo

// Device image-batch allocator: the GPU analogue of the bucketed byte pools in pool.go:57-71.
// One wgpu_ctx per GPU owns the device buffers and the pinned staging memory; Go never holds device pointers.
package pool

/*
#cgo LDFLAGS: -lwebpgpu
#include "webpgpu.h"
*/
import "C"

import (
	"fmt"
	"sync"
	"unsafe"
)

// Device wraps one wgpu_ctx. Calls on one Device are serialised by the library.
type Device struct {
	ctx *C.wgpu_ctx
}

var (
	devMu   sync.Mutex
	devices = map[int]*Device{}
)

// GetDevice returns the process-wide context of GPU `ordinal`, creating it on first use.
// There is no CPU fallback: an error here is returned to the caller of webp.Encode / webp.Decode.
func GetDevice(ordinal int) (*Device, error) {
	devMu.Lock()
	defer devMu.Unlock()
	if d, ok := devices[ordinal]; ok {
		return d, nil
	}
	var ctx *C.wgpu_ctx
	if rc := C.wgpu_ctx_create(C.int(ordinal), &ctx); rc != 0 {
		return nil, fmt.Errorf("webp: cuda: %s", C.GoString(C.wgpu_last_error(nil)))
	}
	d := &Device{ctx: ctx}
	devices[ordinal] = d
	return d, nil
}

// Ctx exposes the handle to the sibling packages (internal/lossy, internal/dsp).
func (d *Device) Ctx() unsafe.Pointer { return unsafe.Pointer(d.ctx) }

// Err formats the library's last error for this context.
func (d *Device) Err(op string) error {
	return fmt.Errorf("webp: cuda: %s: %s", op, C.GoString(C.wgpu_last_error(d.ctx)))
}

// PinnedBytes returns a page-locked staging slice (wgpu_host_alloc); release with FreePinned.
func (d *Device) PinnedBytes(n int) []byte {
	p := C.wgpu_host_alloc(d.ctx, C.size_t(n))
	if p == nil {
		return nil
	}
	return unsafe.Slice((*byte)(p), n)
}

// FreePinned releases a slice obtained from PinnedBytes.
func (d *Device) FreePinned(b []byte) {
	if len(b) > 0 {
		C.wgpu_host_free(d.ctx, unsafe.Pointer(&b[0]))
	}
}

// MemInfo reports what the context's allocator holds: device bytes, pinned host bytes, live buffers (wgpu_ctx_mem_info).
// Buffers are grow-only and sized in the classes of pool.go:14-40 up to 1 MiB, sixteenths of a power of two above.
func (d *Device) MemInfo() (device, pinned uint64, buffers int) {
	var dv, pv C.size_t
	var n C.int
	C.wgpu_ctx_mem_info(d.ctx, &dv, &pv, &n)
	return uint64(dv), uint64(pv), int(n)
}

// Trim returns the working buffers to the driver, as draining the sync.Pool buckets does on the CPU side; constant tables
// stay and the next call re-reserves (wgpu_ctx_trim).
func (d *Device) Trim() error {
	if rc := C.wgpu_ctx_trim(d.ctx); rc != 0 {
		return d.Err("trim")
	}
	return nil
}
