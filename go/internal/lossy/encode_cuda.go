//go:build cuda && cgo

// GPU Phase A of the row-parallel encoder.  Under the cuda tag, EncodeFrame (encode.go:1324) calls
// encodeFrameCUDA instead of encodeFrameParallel (encode_parallel.go:168) when useParallel holds
// (encode.go:1356); analysis(), assignSegments(), setSegmentParams(), recordAllTokens(), optimizeProba()
// and emitFrame() stay exactly as they are.
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

// UseSharpYUV: under the cuda tag webp.encodeLossyWithAlpha (encode.go:531) keeps the RGBA and sets sharp instead of calling
// sharpYUVConvert + NewEncoderFromYUV; the library then derives the source planes itself (sharp_kernels.cuh).
func cudaOptions(cfg *EncodeConfig, sharp ...bool) C.wgpu_enc_options {
	sharpYUV := 0
	if len(sharp) > 0 && sharp[0] {
		sharpYUV = 1
	}
	return C.wgpu_enc_options{
		use_sharp_yuv: C.int(sharpYUV),
		quality: C.int(cfg.Quality), method: C.int(cfg.Method), sns_strength: C.int(cfg.SNSStrength),
		filter_strength: C.int(cfg.FilterStrength), filter_sharpness: C.int(cfg.FilterSharpness),
		filter_type: C.int(cfg.FilterType), partitions: C.int(cfg.Partitions), segments: C.int(cfg.Segments),
		preprocessing: C.int(cfg.Preprocessing), has_alpha: C.int(cfg.HasAlpha),
		passes: C.int(cfg.Pass), dither_amp: C.int(ditherAmp(cfg.Dithering)),
		target_size: C.int(cfg.TargetSize), target_psnr: C.float(cfg.TargetPSNR), qmin: C.int(cfg.QMin), qmax: C.int(cfg.QMax),
	}
}

// ditherAmp mirrors dsp.InitRandom (internal/dsp/random.go:39-50): amp = int(256 * dithering), clamped to [0, 256].
func ditherAmp(d float32) int {
	if d <= 0 {
		return 0
	}
	if d > 1 {
		return 256
	}
	return int(float32(256) * d)
}

func cudaSegQuant(q *SegmentQuant) C.wgpu_seg_quant {
	var out C.wgpu_seg_quant
	out.quant, out.iquant, out.bias = C.int(q.Quant), C.int(q.IQuant), C.int(q.Bias)
	out.dc_quant, out.dc_iquant, out.dc_bias = C.int(q.DCQuant), C.int(q.DCIQuant), C.int(q.DCBias)
	for i := range q.Sharpen {
		out.sharpen[i] = C.int16_t(q.Sharpen[i])
	}
	return out
}

// EncodeBatchCUDA runs import + analysis + mode search for n same-size RGBA images on one GPU and fills each
// encoder's mbInfo / dqm exactly as encodeFrameParallel's Phase A would.  rgba is n images of stride*h bytes.
func EncodeBatchCUDA(encs []*VP8Encoder, rgba []byte, stride int) error {
	dev, err := pool.GetDevice(0)
	if err != nil {
		return err
	}
	if len(encs) == 0 || len(rgba) == 0 {
		return nil
	}
	ctx := (*C.wgpu_ctx)(dev.Ctx())
	n, w, h := len(encs), encs[0].width, encs[0].height
	nmb := encs[0].mbW * encs[0].mbH
	opt := cudaOptions(encs[0].config)
	if rc := C.wgpu_enc_upload(ctx, (*C.uint8_t)(unsafe.Pointer(&rgba[0])), C.int(n), C.int(w), C.int(h), C.int(stride),
		C.size_t(stride*h)); rc != 0 {
		return dev.Err("upload")
	}
	// 1. GPU: importImage + computeAlphas (encode.go:671, encode_analysis.go:245)
	alphas := make([]uint8, n*nmb)
	uvSum := make([]int64, n)
	if rc := C.wgpu_enc_analyze(ctx, &opt, (*C.uint8_t)(&alphas[0]), (*C.int64_t)(&uvSum[0])); rc != 0 {
		return dev.Err("analyze")
	}
	// 2. host, unchanged Go: assignSegments / setSegmentParams / setupSegment / setSegmentProbas
	segs := make([]C.wgpu_segment, n*4)
	segMap := make([]uint8, n*nmb)
	for i, enc := range encs {
		enc.finishAnalysisFromAlphas(alphas[i*nmb:(i+1)*nmb], int(uvSum[i]/int64(nmb))) // encode_analysis.go:310-352 tail
		enc.setSegmentProbas()
		for s := 0; s < 4; s++ {
			d := &enc.dqm[s]
			segs[i*4+s] = C.wgpu_segment{y1: cudaSegQuant(&d.Y1), y2: cudaSegQuant(&d.Y2), uv: cudaSegQuant(&d.UV),
				lambda_i4: C.int(d.LambdaI4), lambda_i16: C.int(d.LambdaI16), lambda_uv: C.int(d.LambdaUV),
				lambda_mode: C.int(d.LambdaMode), tlambda_i4: C.int(d.TLambdaI4), tlambda_i16: C.int(d.TLambdaI16),
				tlambda_sd: C.int(d.TLambdaSD)}
		}
		for k := 0; k < nmb; k++ {
			segMap[i*nmb+k] = enc.mbInfo[k].Segment
		}
	}
	// 3. GPU: pickBestModeParallel + residuals + reconstruction for every MB, wavefront over the batch
	if rc := C.wgpu_enc_search(ctx, &segs[0], (*C.uint8_t)(&segMap[0])); rc != 0 {
		return dev.Err("search")
	}
	// 4. per-MB decisions and levels back into mbInfo (layout of wgpu_enc_fetch == MBEncInfo fields)
	hdr := make([]uint8, nmb*8)
	modes := make([]uint8, nmb*16)
	nz := make([]uint8, nmb*24)
	coeffs := make([]int16, nmb*400)
	for i, enc := range encs {
		if rc := C.wgpu_enc_fetch(ctx, C.int(i), (*C.uint8_t)(&hdr[0]), (*C.uint8_t)(&modes[0]), (*C.uint8_t)(&nz[0]),
			(*C.int16_t)(&coeffs[0]), nil, nil, nil, nil, nil, nil, nil); rc != 0 {
			return dev.Err("fetch")
		}
		for k := 0; k < nmb; k++ {
			info := &enc.mbInfo[k]
			info.MBType, info.I16Mode, info.UVMode = int(hdr[8*k]), hdr[8*k+1], hdr[8*k+2]
			info.Skip, info.NzDC = hdr[8*k+4] != 0, hdr[8*k+5]
			copy(info.Modes[:], modes[16*k:16*k+16])
			copy(info.NzY[:], nz[24*k:24*k+16])
			copy(info.NzUV[:], nz[24*k+16:24*k+24])
			copy(info.Coeffs[:], coeffs[400*k:400*k+400])
		}
	}
	return nil // callers continue with recordAllTokens(&stats) / optimizeProba / emitFrame (encode.go:1370-1400)
}
