//go:build cuda && cgo

// Batched twins of the per-block operator surface (dsp.go:12-37, *Direct functions).  A per-4x4 cgo call is
// meaningless on a GPU, so production code never calls these per block; they exist so the reference's own
// conformance pattern (simd_test.go: backend == scalar Go on random inputs) can be run against the CUDA backend.
package dsp

/*
#cgo LDFLAGS: -lwebpgpu
#include "webpgpu.h"
*/
import "C"

import "unsafe"

// FTransformBatchCUDA: n dense 4x4 tiles of src and ref -> n x 16 coefficients (FTransformDirect, transforms.go:371).
func FTransformBatchCUDA(ctx unsafe.Pointer, src, ref []byte, out []int16) int {
	n := len(out) / 16
	return int(C.wgpu_dsp_ftransform_batch((*C.wgpu_ctx)(ctx), C.int(n), (*C.uint8_t)(&src[0]), (*C.uint8_t)(&ref[0]), (*C.int16_t)(&out[0])))
}

// ITransformBatchCUDA: ITransformDirect (transforms.go:265).
func ITransformBatchCUDA(ctx unsafe.Pointer, ref []byte, in []int16, dst []byte) int {
	n := len(dst) / 16
	return int(C.wgpu_dsp_itransform_batch((*C.wgpu_ctx)(ctx), C.int(n), (*C.uint8_t)(&ref[0]), (*C.int16_t)(&in[0]), (*C.uint8_t)(&dst[0])))
}

// SSE4x4BatchCUDA / TDisto4x4BatchCUDA: SSE4x4Direct, TDisto4x4 (ssim.go:188,315).
func SSE4x4BatchCUDA(ctx unsafe.Pointer, a, b []byte, out []int32) int {
	return int(C.wgpu_dsp_sse4x4_batch((*C.wgpu_ctx)(ctx), C.int(len(out)), (*C.uint8_t)(&a[0]), (*C.uint8_t)(&b[0]), (*C.int32_t)(&out[0])))
}
func TDisto4x4BatchCUDA(ctx unsafe.Pointer, a, b []byte, out []int32) int {
	return int(C.wgpu_dsp_tdisto4x4_batch((*C.wgpu_ctx)(ctx), C.int(len(out)), (*C.uint8_t)(&a[0]), (*C.uint8_t)(&b[0]), (*C.int32_t)(&out[0])))
}

// PredSquareBatchCUDA: PredLuma16Direct / PredChroma8Direct for the seven modes (predict_lossy.go:27-181).
func PredSquareBatchCUDA(ctx unsafe.Pointer, n, size int, ctxPx, out []byte) int {
	return int(C.wgpu_dsp_pred_square_batch((*C.wgpu_ctx)(ctx), C.int(n), C.int(size), (*C.uint8_t)(&ctxPx[0]), (*C.uint8_t)(&out[0])))
}

// UpsampleNRGBACUDA: buildNRGBA / UpsampleLinePairNRGBA over whole planes (webp.go:379, upsample.go:130).
func UpsampleNRGBACUDA(ctx unsafe.Pointer, n, w, h int, y []byte, yStride int, u, v []byte, uvStride, yPlane, uvPlane int, alpha, out []byte) int {
	var pa *C.uint8_t
	if alpha != nil {
		pa = (*C.uint8_t)(&alpha[0])
	}
	return int(C.wgpu_upsample_nrgba((*C.wgpu_ctx)(ctx), C.int(n), C.int(w), C.int(h), (*C.uint8_t)(&y[0]), C.int(yStride), (*C.uint8_t)(&u[0]),
		(*C.uint8_t)(&v[0]), C.int(uvStride), C.size_t(yPlane), C.size_t(uvPlane), pa, (*C.uint8_t)(&out[0])))
}

// PlaneMetricsCUDA: SSE + sum of SSIMGet/SSIMGetClipped per plane pair (ssim.go:116-181).
func PlaneMetricsCUDA(ctx unsafe.Pointer, n int, a, b []byte, w, h, stride, planeStride int, sse []uint64, ssim []float64) int {
	return int(C.wgpu_plane_metrics((*C.wgpu_ctx)(ctx), C.int(n), (*C.uint8_t)(&a[0]), (*C.uint8_t)(&b[0]), C.int(w), C.int(h), C.int(stride),
		C.size_t(planeStride), (*C.uint64_t)(&sse[0]), (*C.double)(&ssim[0])))
}

// CleanupTransparentAreaCUDA: cleanupTransparentAreaLossy (encode.go:788) over n NRGBA images of one size; webp.Encode calls
// it (when !opts.Exact and the image has alpha) before handing the pixels to lossy.NewEncoder.
func CleanupTransparentAreaCUDA(ctx unsafe.Pointer, nrgba []byte, n, w, h, stride, imageStride int, out []byte) int {
	return int(C.wgpu_cleanup_transparent((*C.wgpu_ctx)(ctx), (*C.uint8_t)(&nrgba[0]), C.int(n), C.int(w), C.int(h), C.int(stride),
		C.size_t(imageStride), (*C.uint8_t)(&out[0])))
}

// The rest of the surface (include/webpgpu.h): 16x16 metrics, DequantCoeffs, decoder transforms, the loop-filter set on
// 24x24 tiles, UpsampleLinePair (RGB) / UpsampleLinePairNRGBA.
func SSE16x16BatchCUDA(ctx unsafe.Pointer, a, b []byte, out []int32) int {
	return int(C.wgpu_dsp_sse16x16_batch((*C.wgpu_ctx)(ctx), C.int(len(out)), (*C.uint8_t)(&a[0]), (*C.uint8_t)(&b[0]), (*C.int32_t)(&out[0])))
}
func TDisto16x16BatchCUDA(ctx unsafe.Pointer, a, b []byte, out []int32) int {
	return int(C.wgpu_dsp_tdisto16x16_batch((*C.wgpu_ctx)(ctx), C.int(len(out)), (*C.uint8_t)(&a[0]), (*C.uint8_t)(&b[0]), (*C.int32_t)(&out[0])))
}
func DequantCoeffsBatchCUDA(ctx unsafe.Pointer, in []int16, dcQ, acQ int, out []int16) int {
	return int(C.wgpu_dsp_dequant_batch((*C.wgpu_ctx)(ctx), C.int(len(out)/16), (*C.int16_t)(&in[0]), C.int(dcQ), C.int(acQ), (*C.int16_t)(&out[0])))
}

// kind: 0 Transform, 1 TransformDC, 2 TransformAC3 (4x4 blocks), 3 TransformUV, 4 TransformDCUV (8x8 tiles) -- transforms.go:37-216.
func DecTransformBatchCUDA(ctx unsafe.Pointer, n, kind int, in []int16, ref, dst []byte) int {
	return int(C.wgpu_dsp_dec_transform_batch((*C.wgpu_ctx)(ctx), C.int(n), C.int(kind), (*C.int16_t)(&in[0]), (*C.uint8_t)(&ref[0]), (*C.uint8_t)(&dst[0])))
}

// kind: 0 SimpleVFilter16 ... 11 HFilter8i in the order of filter.go:93-242; tiles are 24x24 with the block at (4, 4).
func FilterBatchCUDA(ctx unsafe.Pointer, n, kind int, tilesIn []byte, thresh, ithresh, hevThresh int, tilesOut []byte) int {
	return int(C.wgpu_dsp_filter_batch((*C.wgpu_ctx)(ctx), C.int(n), C.int(kind), (*C.uint8_t)(&tilesIn[0]), C.int(thresh), C.int(ithresh),
		C.int(hevThresh), (*C.uint8_t)(&tilesOut[0])))
}

// channels = 3: UpsampleLinePair (upsample.go:45); 4: UpsampleLinePairNRGBA (:130).  botY == nil: last row of an odd height.
func UpsampleLinePairBatchCUDA(ctx unsafe.Pointer, n, width int, topY, botY, topU, topV, botU, botV, alphaTop, alphaBot []byte, channels int,
	topDst, botDst []byte) int {
	p := func(b []byte) *C.uint8_t {
		if len(b) == 0 {
			return nil
		}
		return (*C.uint8_t)(&b[0])
	}
	return int(C.wgpu_dsp_upsample_line_pair_batch((*C.wgpu_ctx)(ctx), C.int(n), C.int(width), p(topY), p(botY), p(topU), p(topV), p(botU), p(botV),
		p(alphaTop), p(alphaBot), C.int(channels), p(topDst), p(botDst)))
}

// VP8BitWriter (internal/bitio/writer_bool.go:58-150) over n flat token arrays, bit | prob << 8 per token (encode_token.go:20): the
// encoder's chunk-parallel device coder.  tokens holds the arrays back to back, totals[i] tokens each; partition i is written to
// out[i*outStride:], sizes[i] bytes.  Returns the status and the number of state-relaxation rounds the batch needed.
func BoolCodeBatchCUDA(ctx unsafe.Pointer, tokens []uint16, totals []uint64, out []byte, outStride int, sizes []uint32) (int, int) {
	var rounds C.int
	rc := C.wgpu_dsp_boolcode_batch((*C.wgpu_ctx)(ctx), C.int(len(totals)), (*C.uint16_t)(&tokens[0]), (*C.ulonglong)(&totals[0]),
		(*C.uint8_t)(&out[0]), C.size_t(outStride), (*C.uint)(&sizes[0]), &rounds)
	return int(rc), int(rounds)
}
