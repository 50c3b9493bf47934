// Pure-Go reference number for BASELINE.json's metric (no build tag: runs against the unmodified reference).
// Same synthetic generator as webp_b200/synth.py kind 0/2 (deterministic formulas only, so Go and numpy agree);
//   go test -run xxx -bench LossyEncode1536 -benchtime 20x .
package webp

import (
	"bytes"
	"image"
	"runtime"
	"testing"
)

func synthNoisy(w, h int) *image.RGBA { // race_test.go:78 noisyImage shape without the RNG term
	img := image.NewRGBA(image.Rect(0, 0, w, h))
	for y := 0; y < h; y++ {
		for x := 0; x < w; x++ {
			o := img.PixOffset(x, y)
			img.Pix[o], img.Pix[o+1], img.Pix[o+2], img.Pix[o+3] = uint8((x*7+y*13)%256), uint8((x*3+y*5)%256), uint8((x^y)%256), 255
		}
	}
	return img
}

func BenchmarkLossyEncode1536(b *testing.B) {
	img := synthNoisy(1536, 1024)
	opts := DefaultOptions()
	b.Logf("GOMAXPROCS=%d NumCPU=%d", runtime.GOMAXPROCS(0), runtime.NumCPU())
	b.SetBytes(1536 * 1024) // pixels, so MB/s reads as Mpix/s
	for i := 0; i < b.N; i++ {
		var buf bytes.Buffer
		if err := Encode(&buf, img, opts); err != nil {
			b.Fatal(err)
		}
	}
}
