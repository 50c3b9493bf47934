package lossy

// No build tag: this file goes into the UNMODIFIED reference so that anyone with a Go toolchain can pin the encoder decisions
// the GPU path and its C++ oracle are tested against (SURVEY.md 8c, residual risk ii).  It only reads encoder state.

import (
	"fmt"
	"io"
)

// DumpMBInfo writes one line per macroblock after EncodeFrame: index, type (0 I16 / 1 I4), I16 mode, chroma mode, segment,
// skip, the sixteen 4x4 modes, the nz counts of the 24 blocks and of the WHT block.  tools/refdump_expect.py prints the same
// lines from the oracle (and the GPU) for the same images.
func (enc *VP8Encoder) DumpMBInfo(w io.Writer) {
	for i := range enc.mbInfo {
		m := &enc.mbInfo[i]
		skip := 0
		if m.Skip {
			skip = 1
		}
		fmt.Fprintf(w, "%d %d %d %d %d %d", i, m.MBType, m.I16Mode, m.UVMode, m.Segment, skip)
		for _, v := range m.Modes {
			fmt.Fprintf(w, " %d", v)
		}
		for _, v := range m.NzY {
			fmt.Fprintf(w, " %d", v)
		}
		for _, v := range m.NzUV {
			fmt.Fprintf(w, " %d", v)
		}
		fmt.Fprintf(w, " %d\n", m.NzDC)
	}
}
