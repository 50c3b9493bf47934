// refdump: the reference encoder's decisions and bitstream hash for the RNG-free synthetic pictures of
// webp_b200/synth.py::synth_formula, in the format of tools/refdump_expect.py.  Run inside a checkout of deepteams/webp with
// go/internal/lossy/refdump.go copied to internal/lossy/:
//
//	go run ./cmd/refdump -w 768 -h 576 -kind 1 -q 75 -m 4 > ref.txt
//	python tools/refdump_expect.py 768 576 1 75 4 > ours.txt && diff ref.txt ours.txt
package main

import (
	"crypto/sha256"
	"flag"
	"fmt"
	"image"
	"os"

	"github.com/deepteams/webp/internal/lossy"
)

// synthFormula: kind 0 = richTestImage gradient (root encode_test.go:1496), kind 1 = noisyImage without its random term
// (race_test.go:78), kind 2 = both mixed in 32x32 tiles with hard edges.  Same integer formulas as synth.py::synth_formula.
func synthFormula(w, h, kind int) *image.RGBA {
	img := image.NewRGBA(image.Rect(0, 0, w, h))
	for y := 0; y < h; y++ {
		for x := 0; x < w; x++ {
			var r, g, b int
			k := kind
			if kind == 2 {
				k = ((x >> 5) + (y >> 5)) & 1
			}
			if k == 0 {
				r, g, b = x*255/w, y*255/h, (x+y)*255/(w+h)
			} else {
				r, g, b = (x*7+y*13)%256, (x*3+y*5)%256, (x^y)%256
			}
			o := img.PixOffset(x, y)
			img.Pix[o], img.Pix[o+1], img.Pix[o+2], img.Pix[o+3] = uint8(r), uint8(g), uint8(b), 255
		}
	}
	return img
}

func main() {
	w := flag.Int("w", 768, "width")
	h := flag.Int("h", 576, "height")
	kind := flag.Int("kind", 1, "picture: 0 gradient, 1 noisy formula, 2 tiled mix")
	q := flag.Int("q", 75, "quality")
	m := flag.Int("m", 4, "method")
	flag.Parse()
	cfg := lossy.DefaultConfig(*q) // SNS 50, filter 60 strong, 4 segments, 1 pass: what webp.DefaultOptions maps to
	cfg.Method = *m
	enc := lossy.NewEncoder(synthFormula(*w, *h, *kind), cfg)
	defer lossy.ReleaseEncoder(enc)
	frame, err := enc.EncodeFrame()
	if err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
	fmt.Printf("vp8 %d bytes sha256 %x\n", len(frame), sha256.Sum256(frame))
	enc.DumpMBInfo(os.Stdout)
}
