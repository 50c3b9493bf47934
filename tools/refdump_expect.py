"""What go/cmd/refdump prints, from this repo: the VP8 frame's size + SHA-256 and one line per macroblock (type, I16 mode,
chroma mode, segment, skip, sixteen 4x4 modes, 24 nz counts, WHT nz count) for an RNG-free synthetic picture.
  python tools/refdump_expect.py W H KIND QUALITY METHOD [--gpu]
Default source: the C++ oracle (oracle/); --gpu takes the same fields from the CUDA path (wgpu_enc_fetch) and checks that both
agree.  Diff the output against `go run ./cmd/refdump -w W -h H -kind KIND -q QUALITY -m METHOD` in a deepteams/webp checkout
to pin the encoder decisions against the Go reference."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from webp_b200.synth import synth_formula

w, h, kind, q, m = [int(a) for a in sys.argv[1:6]]
img = synth_formula(w, h, kind)
data, t = oracle_lib.encode(img, oracle_lib.default_cfg(quality=q, method=m), taps=True)
frame = data[20:20 + int.from_bytes(data[16:20], "little")]  # payload of the "VP8 " chunk of the simple RIFF file
lines = ["vp8 %d bytes sha256 %s" % (len(frame), hashlib.sha256(frame).hexdigest())]
for i in range(len(t["mb_hdr"])):
    hd = t["mb_hdr"][i]
    lines.append(" ".join(str(int(v)) for v in [i, hd[0], hd[1], hd[2], hd[3], hd[4], *t["mb_modes"][i], *t["mb_nz"][i], hd[5]]))
if "--gpu" in sys.argv:
    import webp_b200
    from webp_b200 import native
    o = webp_b200.DefaultOptions(); o.Quality = q; o.Method = m
    ctx = native.Context(0)
    assert webp_b200.EncodeBatch(img[None], o, ctx)[0] == data, "GPU bytes differ from the oracle"
print("\n".join(lines))
