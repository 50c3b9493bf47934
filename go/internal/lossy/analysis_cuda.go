//go:build cuda && cgo

package lossy

// finishAnalysisFromAlphas is analysis() (encode_analysis.go:29-74) without its first step: the per-macroblock alphas and the
// global chroma alpha come from the GPU (wgpu_enc_analyze == computeAlphas, encode_analysis.go:245-353), everything after them
// -- k-means segment assignment, setSegmentParams, the segment header -- is the reference's own code, called unchanged.
func (enc *VP8Encoder) finishAnalysisFromAlphas(gpuAlphas []uint8, globalUVAlpha int) {
	numSegs := enc.config.Segments
	if numSegs < 1 {
		numSegs = 1
	}
	if numSegs > NumMBSegments {
		numSegs = NumMBSegments
	}
	alphas := enc.analysisAlphas[:len(enc.mbInfo)]
	for i := range alphas {
		alphas[i] = int(gpuAlphas[i])
		enc.mbInfo[i].Alpha = alphas[i] // computeAlphas stores the mixed alpha here too (encode_analysis.go:300)
	}
	enc.globalAlpha = 0
	enc.globalUVAlpha = globalUVAlpha
	if numSegs <= 1 {
		for i := range enc.mbInfo {
			enc.mbInfo[i].Segment = 0
		}
		enc.dqm[0].Alpha = 0
		enc.dqm[0].Beta = 0
	} else {
		assignSegments(enc, alphas, numSegs)
	}
	enc.setSegmentParams(numSegs)
	enc.buildSegmentHeader(enc.numSegments)
}

// finishFrameCUDA is the tail of EncodeFrame (encode.go:1370-1400) for the row-parallel path once mbInfo holds the GPU's
// decisions: Phase B of encodeFrameParallel (token recording + statistics, encode_parallel.go:237-243), the final probability
// optimisation and the bitstream.  enc.parallelRS stays nil: every row is already complete.
func (enc *VP8Encoder) finishFrameCUDA() ([]byte, error) {
	var stats ProbaStats
	enc.tokens.Reset()
	enc.recordAllTokens(&stats)
	if optimizeProba(&stats, &enc.proba) > 0 {
		enc.rerecordAllTokens()
	}
	frameData, err := enc.emitFrame()
	if err != nil {
		return nil, err
	}
	enc.computeStats(frameData)
	return frameData, nil
}

// EncodeFramesCUDA is EncodeFrame for a batch of same-size encoders on the row-parallel path (Method >= 3, no rate control,
// at least 4 macroblock rows -- the useParallel condition of encode.go:1356; anything else keeps the reference's own path).
func EncodeFramesCUDA(encs []*VP8Encoder, rgba []byte, stride int) ([][]byte, error) {
	if len(encs) == 0 {
		return nil, nil
	}
	if err := EncodeBatchCUDA(encs, rgba, stride); err != nil {
		return nil, err
	}
	out := make([][]byte, len(encs))
	for i, enc := range encs {
		data, err := enc.finishFrameCUDA()
		if err != nil {
			return nil, err
		}
		out[i] = data
	}
	return out, nil
}
