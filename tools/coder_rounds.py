"""How many state-relaxation rounds the chunk-parallel boolean coder needs: real token partitions vs adversarial streams.
  python tools/coder_rounds.py   (needs a GPU; test infrastructure: compares with the oracle's writer)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from webp_b200 import native, dsp
ctx = native.Context(0)
rng = np.random.default_rng(11)
cases = {"real 640x480 noisy": [O.encode_tokens(O.synth_image(640, 480, 2, kind=2))[0]],
         "real 512x512 smooth": [O.encode_tokens(O.synth_image(512, 512, 3, kind=0))[0]]}
for m in range(6):
    cases["adversarial mode %d, 300 k tokens" % m] = [O.adversarial_tokens(rng, 300000, m)]
for name, streams in cases.items():
    got, rounds = dsp.BoolCodeBatch(streams, ctx)
    ok = all(np.array_equal(g, O.boolcode(t)) for g, t in zip(got, streams))
    print("%-36s %8d tokens  rounds %4d  %s" % (name, len(streams[0]), rounds, "ok" if ok else "MISMATCH"))
