//go:build cuda && cgo

// GPU reconstruction + loop filter.  Under the cuda tag DecodeFrame (decode.go:209) keeps parseHeaders /
// parseIntraModeRow / decodeMB (bool decoder, host) but stores every row's MBData + FInfo in a frame-sized
// array instead of one row (decode.go:463-468), then hands the whole frame -- or a batch of frames -- to the GPU
// in place of reconstructRow + filterRowAt (decode_frame.go:83-342).
package lossy

/*
#cgo LDFLAGS: -lwebpgpu
#include "webpgpu.h"
*/
import "C"

import (
	"unsafe"

	"github.com/deepteams/webp/internal/pool"
)

// DecodeBatchCUDA decodes n lossy streams of identical dimensions.  y/u/v receive macroblock-padded planes
// (stride 16*mbW / 8*mbW, as dec.cacheY/U/V); nrgba, when non-nil, receives buildNRGBA output (webp.go:379).
// The library restates the host parser; the Go parser can be used instead through wgpu_dec_* once a
// "pre-parsed MBData" entry point is added (SURVEY.md 8b) -- both produce the same MBData by construction.
func DecodeBatchCUDA(streams [][]byte, y, u, v, nrgba []byte, yPlane, uvPlane, nrgbaImage int) error {
	dev, err := pool.GetDevice(0)
	if err != nil {
		return err
	}
	ctx := (*C.wgpu_ctx)(dev.Ctx())
	n := len(streams)
	ptrs := (**C.uint8_t)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0))))) // C array: no Go pointers to Go pointers
	lens := make([]C.size_t, n)
	defer C.free(unsafe.Pointer(ptrs))
	pp := unsafe.Slice(ptrs, n)
	for i, s := range streams {
		pp[i] = (*C.uint8_t)(C.CBytes(s))
		lens[i] = C.size_t(len(s))
		defer C.free(unsafe.Pointer(pp[i]))
	}
	var pn *C.uint8_t
	if nrgba != nil {
		pn = (*C.uint8_t)(&nrgba[0])
	}
	if rc := C.wgpu_decode_batch(ctx, ptrs, &lens[0], C.int(n), (*C.uint8_t)(&y[0]), (*C.uint8_t)(&u[0]), (*C.uint8_t)(&v[0]),
		C.size_t(yPlane), C.size_t(uvPlane), pn, C.size_t(nrgbaImage)); rc != 0 {
		return dev.Err("decode")
	}
	return nil
}
