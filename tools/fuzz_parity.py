"""Randomised differential test: random sizes x random EncoderOptions, GPU (through the public Python mirror / C ABI) vs the
oracle, bytes and decoded planes.  python tools/fuzz_parity.py [cases] [seed]   (needs a GPU; test infrastructure)"""
import os, sys, random, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import webp_b200
from webp_b200 import native



def fuzz(cases, seed, ctx=None):
    """Returns the list of failing case descriptions."""
    rnd = random.Random(seed)
    ctx = ctx or native.Context(0)
    bad = 0
    failures = []
    saved = {k: os.environ.get(k) for k in ("WGPU_DEVICE_CODER", "WGPU_DEVICE_PARSER")}
    for c in range(cases):
        w, h = rnd.choice([(16, 16), (33, 17), (64, 48), (100, 70), (128, 96), (130, 71), (200, 150), (256, 64), (320, 240), (48, 200), (640, 40)])
        if rnd.random() < 0.3:
            w, h = rnd.randint(1, 260), rnd.randint(1, 200)
        if os.environ.get("FUZZ_BIG") and rnd.random() < 0.5:  # several probability refreshes on the serial paths
            w, h = rnd.choice([(400, 300), (640, 360), (512, 512), (1000, 48), (333, 777)])
        o = webp_b200.DefaultOptions()
        o.Quality = rnd.choice([0, 5, 20, 40, 49, 50, 60, 75, 90, 98, 100])
        o.Method = rnd.randint(0, 6)
        o.SNSStrength = rnd.choice([-1, 0, 30, 50, 100])
        o.FilterStrength = rnd.choice([-1, 0, 20, 60, 100])
        o.FilterSharpness = rnd.choice([0, 0, 3, 7])
        o.FilterType = rnd.choice([-1, 0, 1])
        o.Segments = rnd.choice([-1, 1, 2, 3, 4])
        o.Preprocessing = rnd.choice([0, 0, 1, 2, 3])
        o.Pass = rnd.choice([-1, 1, 2, 4])
        mode = rnd.random()
        nmb = ((w + 15) // 16) * ((h + 15) // 16)
        serial = o.Method >= 3 and (((h + 15) // 16) < 4)
        if mode < 0.25:
            o.TargetSize = rnd.choice([300, 1000, 3000, 8000]); serial = True
        elif mode < 0.45:
            o.TargetPSNR = rnd.choice([30.0, 38.5, 45.0]); serial = True
            o.QMin, o.QMax = rnd.choice([(0, -1), (10, 80), (0, 100), (30, 60)])
        o.Partitions = rnd.choice([0, 0, 0, 1, 2, 3])
        if (serial and nmb > 96) or o.TargetSize > 0 or o.TargetPSNR > 0:
            o.Partitions = 0  # the refresh route and rate control are single-partition
        idxs = [rnd.randint(0, 11) for _ in range(rnd.choice([1, 2, 3]))]
        o.UseSharpYUV = random.Random(seed * 7919 + c).random() < 0.2  # own generator: the other draws of a (seed, case) stay what they were
        os.environ["WGPU_DEVICE_CODER"] = rnd.choice(["0", "1"])
        os.environ["WGPU_DEVICE_PARSER"] = rnd.choice(["0", "1"])
        desc = "case %d: %dx%d imgs=%s coder=%s parser=%s %s" % (c, w, h, idxs, os.environ["WGPU_DEVICE_CODER"], os.environ["WGPU_DEVICE_PARSER"],
                                                                 {k: v for k, v in vars(o).items() if v != getattr(webp_b200.DefaultOptions(), k)})
        try:
            imgs = np.stack([O.synth_image(w, h, i) for i in idxs])
            files = webp_b200.EncodeBatch(imgs, o, ctx)
            cfg = webp_b200.webp.lossy_config(o)
            ocfg = O.default_cfg(**{f: getattr(cfg, f) for f, _ in cfg._fields_})
            exp = [O.encode(imgs[k], ocfg) for k in range(len(idxs))]
            ok = all(files[k] == exp[k] for k in range(len(idxs)))
            decodable = True
            for k in range(len(idxs)):
                try:
                    O.decode(files[k])
                except RuntimeError:
                    decodable = False  # multi-partition streams with skipped macroblocks are corrupt in the reference itself
            if ok and not decodable:
                try:
                    webp_b200.webp.decode_padded(files, nrgba=True, ctx=ctx)
                    ok = False  # the GPU decoder must reject what the oracle decoder rejects
                except (native.WebPGPUError, webp_b200.WebPError):
                    pass
            elif ok:
                gw, gh, y, u, v, rgba = webp_b200.webp.decode_padded(files, nrgba=True, ctx=ctx)
                for k in range(len(idxs)):
                    _, _, ey, eu, ev = O.decode(files[k])
                    ok &= bool(np.array_equal(y[k], ey) and np.array_equal(u[k], eu) and np.array_equal(v[k], ev) and
                               np.array_equal(rgba[k], O.build_nrgba(w, h, ey, eu, ev)))
            if not ok:
                bad += 1
                failures.append("MISMATCH " + desc)
                print("MISMATCH", desc, [(len(a), len(b)) for a, b in zip(files, exp)])
        except Exception as e:  # noqa
            bad += 1
            failures.append("ERROR " + desc + " " + repr(e)[:200])
            print("ERROR", desc, repr(e)[:300])
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return failures


def fuzz_foreign(cases, seed, ctx=None):
    """Streams of a foreign encoder (libwebp through Pillow: loop-filter deltas, quantiser deltas, modes and segment maps our
    encoder never emits) through both macroblock-parser routes, planes against the oracle decoder."""
    import io
    from PIL import Image
    rnd = random.Random(seed)
    ctx = ctx or native.Context(0)
    failures = []
    saved = os.environ.get("WGPU_DEVICE_PARSER")
    for c in range(cases):
        w, h = rnd.randint(1, 300), rnd.randint(1, 220)
        q, m = rnd.choice([1, 10, 30, 50, 75, 90, 100]), rnd.randint(0, 6)
        n = rnd.choice([1, 2, 3])
        streams = []
        for k in range(n):
            buf = io.BytesIO()
            Image.fromarray(O.synth_image(w, h, rnd.randint(0, 11))[..., :3]).save(buf, "WEBP", quality=q, method=m)
            streams.append(buf.getvalue())
        os.environ["WGPU_DEVICE_PARSER"] = rnd.choice(["0", "1"])
        desc = "foreign case %d: %dx%d q%d m%d n=%d parser=%s" % (c, w, h, q, m, n, os.environ["WGPU_DEVICE_PARSER"])
        try:
            gw, gh, y, u, v, rgba = webp_b200.webp.decode_padded(streams, nrgba=True, ctx=ctx)
            for k in range(n):
                _, _, ey, eu, ev = O.decode(streams[k])
                if not (np.array_equal(y[k], ey) and np.array_equal(u[k], eu) and np.array_equal(v[k], ev) and
                        np.array_equal(rgba[k], O.build_nrgba(w, h, ey, eu, ev))):
                    failures.append("MISMATCH " + desc)
                    print("MISMATCH", desc)
                    break
        except Exception as e:  # noqa
            failures.append("ERROR " + desc + " " + repr(e)[:200])
            print("ERROR", desc, repr(e)[:300])
    if saved is None:
        os.environ.pop("WGPU_DEVICE_PARSER", None)
    else:
        os.environ["WGPU_DEVICE_PARSER"] = saved
    return failures


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    the_seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    fails = fuzz(n_cases, the_seed)
    fails += fuzz_foreign(max(10, n_cases // 3), the_seed)
    print("fuzz: %d cases, %d bad (seed %d)" % (n_cases, len(fails), the_seed))
    sys.exit(1 if fails else 0)
