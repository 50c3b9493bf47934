//go:build cuda && cgo

// Batch entry points of the GPU build.  webp.Encode / webp.Decode keep their signatures and behaviour (encode.go:424,
// webp.go:88); one image cannot fill a B200, so the GPU path is reached through the batch twins below, which produce exactly
// the bytes / pixels n calls of Encode / Decode would.
package webp

import (
	"fmt"
	"image"
	"image/draw"
	"io"

	"github.com/deepteams/webp/internal/container"
	"github.com/deepteams/webp/internal/lossy"
)

// lossyConfigCUDA is the EncoderOptions -> lossy.EncodeConfig mapping of encodeLossyWithAlpha (encode.go:477-526), which the
// reference spells inline there; a maintainer would hoist that block into a function and call it from both places.
func lossyConfigCUDA(opts *EncoderOptions) lossy.EncodeConfig {
	cfg := lossy.DefaultConfig(int(opts.Quality))
	cfg.Method = opts.Method
	cfg.QMin = opts.QMin
	cfg.QMax = resolveQMax(opts.QMax)
	if opts.SNSStrength >= 0 {
		cfg.SNSStrength = opts.SNSStrength
	}
	if opts.FilterStrength >= 0 {
		cfg.FilterStrength = opts.FilterStrength
	}
	cfg.FilterSharpness = opts.FilterSharpness
	if opts.FilterType >= 0 {
		cfg.FilterType = opts.FilterType
	}
	cfg.Partitions = opts.Partitions
	if opts.Segments > 0 {
		cfg.Segments = opts.Segments
	}
	if opts.Pass > 0 {
		cfg.Pass = opts.Pass
	}
	cfg.Preprocessing = opts.Preprocessing
	if opts.Preprocessing&2 != 0 {
		x := opts.Quality / 100.0
		x2 := x * x
		cfg.Dithering = 1.0 + (0.5-1.0)*x2*x2
	}
	cfg.HasAlpha = 0
	return cfg
}

// EncodeBatch encodes same-size opaque images with one set of options and writes file i to ws[i].  Options outside the GPU
// lossy path (Lossless, images with alpha, Method < 3 or rate control -- which the library serves through
// wgpu_encode_batch rather than through the mbInfo hand-over used here) fall back to Encode, image by image.
func EncodeBatch(ws []io.Writer, imgs []image.Image, opts *EncoderOptions) error {
	if len(ws) != len(imgs) {
		return fmt.Errorf("webp: EncodeBatch: %d writers for %d images", len(ws), len(imgs))
	}
	if len(imgs) == 0 {
		return nil
	}
	if opts == nil {
		opts = DefaultOptions()
	}
	b := imgs[0].Bounds()
	gpu := !opts.Lossless && !opts.UseSharpYUV && opts.Method >= 3 && opts.TargetSize == 0 && opts.TargetPSNR == 0 && (b.Dy()+15)/16 >= 4
	for _, im := range imgs {
		if im == nil {
			return fmt.Errorf("webp: nil image")
		}
		if im.Bounds().Dx() != b.Dx() || im.Bounds().Dy() != b.Dy() || !isOpaque(im) {
			gpu = false
		}
	}
	if !gpu {
		for i, im := range imgs {
			if err := Encode(ws[i], im, opts); err != nil {
				return err
			}
		}
		return nil
	}
	w, h := b.Dx(), b.Dy()
	stride := 4 * w
	rgba := make([]byte, len(imgs)*stride*h)
	encs := make([]*lossy.VP8Encoder, len(imgs))
	cfg := lossyConfigCUDA(opts)
	for i, im := range imgs {
		dst := &image.RGBA{Pix: rgba[i*stride*h : (i+1)*stride*h], Stride: stride, Rect: image.Rect(0, 0, w, h)}
		draw.Draw(dst, dst.Rect, im, im.Bounds().Min, draw.Src)
		encs[i] = lossy.NewEncoder(dst, cfg) // its own importImage result is replaced by the GPU's (same arithmetic)
		defer lossy.ReleaseEncoder(encs[i])
	}
	frames, err := lossy.EncodeFramesCUDA(encs, rgba, stride)
	if err != nil {
		return fmt.Errorf("webp: encoding VP8 (cuda): %w", err)
	}
	for i, f := range frames {
		if err := writeRIFF(ws[i], container.FourCCVP8, f, nil, w, h, opts); err != nil { // encode.go:955
			return err
		}
	}
	return nil
}

func isOpaque(im image.Image) bool {
	if o, ok := im.(interface{ Opaque() bool }); ok {
		return o.Opaque()
	}
	return false
}
