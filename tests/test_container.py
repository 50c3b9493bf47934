"""Container / animation frame path (webp_b200/mux.py, webp_b200/animation.py).

CPU suite: the host logic driven with the ORACLE as the frame codec; every file it writes is read back by libwebp 1.6.0
(Pillow's WebPAnimDecoder: an independent demuxer + compositor) and must give the canvases our own AnimDecoder composites from
oracle-decoded frames.  GPU suite: the same encoder with the GPU codec must write the same bytes, the batched paths enter the
codec once, and DecodeFrames equals the oracle's planes through ycbcrToNRGBA."""
import io
import struct

import numpy as np
import pytest

import webp_b200
from webp_b200 import animation, mux


def _ocfg(oracle, o):
    c = webp_b200.webp.lossy_config(o)
    return oracle.default_cfg(**{f: getattr(c, f) for f, _ in c._fields_})


def _oracle_codec(oracle, quality):
    """encodeFrameForAnimation (webp.go:212) with the oracle as the codec: EncoderOptions{Quality, Method: 4}, raw VP8 out."""
    cfg = _ocfg(oracle, webp_b200.EncoderOptions(Quality=float(quality), Method=4))
    return lambda frames: [mux.riff_payload(oracle.encode(f, cfg)) for f in frames]


def _oracle_simple(oracle, quality):
    cfg = _ocfg(oracle, webp_b200.EncoderOptions(Quality=float(quality), Method=4))
    return lambda canvas: oracle.encode(canvas, cfg)


def _clip(w, h, n, seed=1):
    """Opaque frames: a textured background, a box that moves, a few repeated frames and one full-canvas change."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 255 // max(w + h - 2, 1)), np.full_like(xx, 255)], axis=2).astype(np.uint8)
    frames = []
    for i in range(n):
        f = base.copy()
        if i == n - 2:
            f[..., :3] = 255 - base[..., :3]  # changes everything: the keyframe candidate
        elif i % 4 != 3:  # every fourth frame repeats its predecessor
            x0, y0 = (3 + 5 * (i % 7)) % max(w - w // 4, 1), (1 + 3 * (i % 5)) % max(h - h // 3, 1)
            f[y0:y0 + h // 3, x0:x0 + w // 4, :3] = rng.integers(0, 256, (h // 3, w // 4, 3), dtype=np.uint8)
        elif frames:
            f = frames[-1].copy()
        frames.append(f)
    return frames


def _fancy_frames(oracle, anim):
    """Frames decoded as libwebp shows them (fancy upsampling): the libwebp cross-check's view of the same file."""
    for f in anim.Frames:
        w, h, y, u, v = oracle.decode(mux.writeRIFFSimple(mux.FourCCVP8, f.BitstreamData))
        f.Image = oracle.build_nrgba(w, h, y, u, v)


def _pillow_canvases(data):
    from PIL import Image
    im = Image.open(io.BytesIO(data))
    out = []
    for i in range(getattr(im, "n_frames", 1)):
        im.seek(i)
        out.append((np.asarray(im.convert("RGBA")).copy(), im.info.get("duration", 0)))
    return out, im.info


def test_riff_writers_spelled_out():
    """encode.go:955-1121: chunk order RIFF, VP8X, ICCP, VP8, EXIF, XMP; flag bits; odd payloads padded; sizes."""
    bs = b"\x10\x02\x00\x9d\x01\x2a\x05\x00\x03\x00" + b"abc"  # 13 bytes: odd
    simple = mux.writeRIFFSimple(mux.FourCCVP8, bs)
    assert simple == b"RIFF" + struct.pack("<I", 4 + 8 + 14) + b"WEBP" + b"VP8 " + struct.pack("<I", 13) + bs + b"\0"
    assert mux.riff_payload(simple) == bs
    ext = mux.writeRIFF(mux.FourCCVP8, bs, None, 5, 3, icc=b"ICC", exif=b"EX", xmp=b"")
    exp = b"VP8X" + struct.pack("<I", 10) + bytes([0x20 | 0x08, 0, 0, 0]) + bytes([4, 0, 0, 2, 0, 0])
    exp += b"ICCP" + struct.pack("<I", 3) + b"ICC\0" + b"VP8 " + struct.pack("<I", 13) + bs + b"\0" + b"EXIF" + struct.pack("<I", 2) + b"EX"
    assert ext == b"RIFF" + struct.pack("<I", 4 + len(exp)) + b"WEBP" + exp
    assert mux.writeRIFF(mux.FourCCVP8, bs, None, 5, 3) == simple
    d = mux.Demuxer(ext)
    assert (d.Width, d.Height, d.HasICC, d.HasEXIF, d.HasXMP, d.HasAnimation, d.HasAlpha) == (5, 3, True, True, False, False, False)
    assert d.iccData == b"ICC" and d.exifData == b"EX" and d.Frame(0).Data == bs and d.NumFrames() == 1


def test_muxer_anmf_layout_and_demux_round_trip():
    """mux/mux.go:331-560: VP8X (animation flag, canvas - 1), ANIM (background, loops), ANMF (offsets / 2, size - 1, duration,
    dispose | blend << 1), frames back out of the Demuxer with the same options."""
    f0 = b"\x10\x02\x00\x9d\x01\x2a\x08\x00\x06\x00" + b"x"      # 8 x 6, odd length
    f1 = b"\x10\x02\x00\x9d\x01\x2a\x04\x00\x02\x00" + b"yz"     # 4 x 2
    m = mux.Muxer()
    m.SetCanvasSize(8, 6); m.SetLoopCount(3); m.SetBackgroundColor(0x11223344)
    m.AddFrame(f0, mux.FrameOptions(Duration=40, BlendMode=mux.BlendNone))
    m.AddFrame(f1, mux.FrameOptions(Duration=1 << 30, OffsetX=4, OffsetY=2, BlendMode=mux.BlendAlpha, DisposeMode=mux.DisposeBackground))
    m.SetXMP(b"<x/>")
    data = m.Assemble()
    body = b"VP8X" + struct.pack("<I", 10) + bytes([0x02 | 0x04, 0, 0, 0, 7, 0, 0, 5, 0, 0])
    body += b"ANIM" + struct.pack("<IIH", 6, 0x11223344, 3)
    body += b"ANMF" + struct.pack("<I", 16 + 8 + 12) + bytes([0, 0, 0, 0, 0, 0, 7, 0, 0, 5, 0, 0, 40, 0, 0, 2]) + b"VP8 " + struct.pack("<I", 11) + f0 + b"\0"
    body += b"ANMF" + struct.pack("<I", 16 + 8 + 12) + bytes([2, 0, 0, 1, 0, 0, 3, 0, 0, 1, 0, 0, 0xff, 0xff, 0xff, 1]) + b"VP8 " + struct.pack("<I", 12) + f1
    body += b"XMP " + struct.pack("<I", 4) + b"<x/>"
    assert data == b"RIFF" + struct.pack("<I", 4 + len(body)) + b"WEBP" + body
    d = mux.Demuxer(data)
    assert d.HasAnimation and d.HasXMP and not d.HasAlpha and (d.Width, d.Height, d.LoopCount(), d.BackgroundColor()) == (8, 6, 3, 0x11223344)
    a, b = d.Frame(0), d.Frame(1)
    assert (a.Data, a.Duration, a.OffsetX, a.OffsetY, a.BlendMode, a.DisposeMode, a.IsKeyframe) == (f0, 40, 0, 0, mux.BlendNone, mux.DisposeNone, True)
    assert (b.Data, b.Duration, b.OffsetX, b.OffsetY, b.BlendMode, b.DisposeMode, b.Width, b.Height) == (f1, 0xFFFFFF, 4, 2, mux.BlendAlpha, mux.DisposeBackground, 4, 2)
    with pytest.raises(mux.MuxError):  # frame outside the canvas (mux.go:233)
        bad = mux.Muxer(); bad.SetCanvasSize(8, 6); bad.AddFrame(f0, mux.FrameOptions(Duration=1, OffsetX=2)); bad.Assemble()
    with pytest.raises(mux.MuxError):
        mux.Muxer().Assemble()


def _vp8_keyframe(w, h):
    """makeVP8Keyframe (mux/mux_test.go:104): frame tag, signature, dimensions."""
    return b"\0\0\0\x9d\x01\x2a" + struct.pack("<HH", w, h)


def test_reference_mux_cases():
    """The cases the reference's own mux tests hold (mux/mux_test.go), transcribed: clamping of durations (:799-880) and loop
    counts (:822), canvas size rules (:934-1010), demux of the two decode fixtures and of a hand-built extended file (:240),
    the empty / missing frame errors (:610-625), bad RIFF (:409), frame index out of range (:416)."""
    for dur, want in ((-1, 0), (0, 0), (100, 100), (0xFFFFFF, 0xFFFFFF), (0x1000000, 0xFFFFFF), (0x7FFFFFFF, 0xFFFFFF), (-50, 0)):
        m = mux.Muxer()
        m.AddFrame(_vp8_keyframe(10, 10), mux.FrameOptions(Duration=dur))
        assert m.FrameDuration(0) == want
        m.SetFrameDuration(0, dur)
        assert m.FrameDuration(0) == want
    for loops, want in ((-1, 0), (0, 0), (3, 3), (0xFFFF, 0xFFFF), (0x10000, 0xFFFF), (0x7FFFFFFF, 0xFFFF)):
        m = mux.Muxer()
        m.SetLoopCount(loops)
        assert m.loopCount == want
        m.AddFrame(_vp8_keyframe(4, 4), mux.FrameOptions(Duration=10))
        assert mux.Demuxer(m.Assemble()).LoopCount() == want  # TestLoopCountRoundtrip (:883)
    m = mux.Muxer(); m.SetICCProfile(b"\0"); m.SetCanvasSize(200, 160); m.AddFrame(_vp8_keyframe(100, 80))
    d = mux.Demuxer(m.Assemble())
    assert (d.Width, d.Height) == (200, 160)  # the explicit canvas wins (:934)
    m = mux.Muxer(); m.SetICCProfile(b"\0"); m.AddFrame(_vp8_keyframe(100, 80))
    d = mux.Demuxer(m.Assemble())
    assert (d.Width, d.Height) == (100, 80) and d.HasICC and d.iccData == b"\0"  # fallback: the frame extent (:960)
    m = mux.Muxer(); m.SetCanvasSize(50, 40)
    assert m.canvasSize() == (50, 40)  # (:987)
    m = mux.Muxer(); m.SetCanvasSize(100, 100)
    m.AddFrame(_vp8_keyframe(50, 50), mux.FrameOptions(Duration=100))
    m.AddFrame(_vp8_keyframe(50, 50), mux.FrameOptions(Duration=100, OffsetX=50, OffsetY=50))
    d = mux.Demuxer(m.Assemble())
    assert (d.Width, d.Height, d.NumFrames()) == (100, 100, 2) and (d.Frame(1).OffsetX, d.Frame(1).OffsetY) == (50, 50)  # (:1000)
    assert d.Frame(0).IsKeyframe and not d.Frame(1).IsKeyframe
    with pytest.raises(mux.MuxError):
        d.Frame(2)
    with pytest.raises(mux.MuxError):
        mux.Muxer().AddFrame(b"")
    with pytest.raises(mux.MuxError):
        mux.Demuxer(b"not a riff file at all")
    import os
    data_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
    for name, dims in (("blue_16x16_lossy.webp", (16, 16)), ("red_4x4_lossy.webp", (4, 4))):  # the reference's decode fixtures (webp_test.go:148-188)
        raw = open(os.path.join(data_dir, name), "rb").read()
        d = mux.Demuxer(raw)
        assert (d.Width, d.Height, d.NumFrames(), d.Format, d.HasAlpha, d.HasAnimation) == dims + (1, 1, False, False)
        assert mux.writeRIFFSimple(mux.FourCCVP8, d.Frame(0).Data) == raw[:8 + struct.unpack_from("<I", raw, 4)[0]]
        m = mux.Muxer(); m.AddFrame(d.Frame(0).Data)
        assert m.Assemble() == mux.writeRIFFSimple(mux.FourCCVP8, d.Frame(0).Data)  # one frame, no metadata: the simple format (:437)
    # VP8L frames pass through as opaque byte strings; their header's alpha bit raises the VP8X alpha flag (:691-800)
    vp8l = bytes([0x2F]) + struct.pack("<I", (9) | (9 << 14) | (1 << 28))
    m = mux.Muxer(); m.AddFrame(vp8l, mux.FrameOptions(Duration=5))
    d = mux.Demuxer(m.Assemble())
    assert d.HasAlpha and d.HasAnimation and d.Frame(0).HasAlpha and (d.Frame(0).Width, d.Frame(0).Height) == (10, 10)


def test_keyframe_options_and_rect_helpers():
    """animation/animation.go:546 (sanitizeKeyframeOptions), :1019 findChangedRect against the reference's scan, :1099 snapToEven."""
    big = (1 << 63) - 1
    assert animation.sanitizeKeyframeOptions(0, 0) == (big - 1, big)
    assert animation.sanitizeKeyframeOptions(5, 1) == (0, 0)
    assert animation.sanitizeKeyframeOptions(9, 4) == (3, 4)
    assert animation.sanitizeKeyframeOptions(1, 10) == (6, 10)
    assert animation.sanitizeKeyframeOptions(1, 100) == (70, 100)
    assert animation.snapToEven((3, 5, 9, 6)) == (2, 4, 9, 6) and animation.snapToEven((2, 4, 3, 5)) == (2, 4, 3, 5)
    assert [animation.qualityToMaxDiff(q) for q in (0, 25, 75, 100)] == [31, 16, 5, 1]
    rng = np.random.default_rng(5)

    def scan(prev, curr):  # the reference's row scan with progressive narrowing, literally
        h, w = prev.shape[:2]
        rows = [y for y in range(h) if not np.array_equal(prev[y], curr[y])]
        if not rows:
            return (0, 0, 0, 0)
        minY, maxY = rows[0], rows[-1] + 1
        minX, maxX = w, 0
        for y in range(minY, maxY):
            for x in range(0, minX):
                if not np.array_equal(prev[y, x], curr[y, x]):
                    minX = x
                    break
            for x in range(w - 1, maxX - 1, -1):
                if not np.array_equal(prev[y, x], curr[y, x]):
                    maxX = x + 1
                    break
        return (minX, minY, maxX, maxY) if maxX > minX else (0, 0, 0, 0)
    for _ in range(40):
        h, w = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        prev = rng.integers(0, 4, (h, w, 4), dtype=np.uint8)
        curr = prev.copy()
        for _ in range(int(rng.integers(0, 4))):
            curr[int(rng.integers(0, h)), int(rng.integers(0, w)), int(rng.integers(0, 4))] ^= 1
        assert animation.findChangedRect(prev, curr) == scan(prev, curr)


def test_ycbcr_to_nrgba_matches_the_scalar_formula():
    """webp.go:272-325 restated per pixel in Python ints (Go's >> on a negative int32 is arithmetic, as Python's)."""
    rng = np.random.default_rng(2)
    h, w = 5, 7
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    cb = rng.integers(0, 256, (3, 4), dtype=np.uint8)
    cr = rng.integers(0, 256, (3, 4), dtype=np.uint8)
    got = animation.ycbcrToNRGBA(y, cb, cr)
    for j in range(h):
        for i in range(w):
            yy, b, r = int(y[j, i]), int(cb[j >> 1, i >> 1]) - 128, int(cr[j >> 1, i >> 1]) - 128
            exp = [yy + ((91881 * r + 32768) >> 16), yy - ((22554 * b + 46802 * r + 32768) >> 16), yy + ((116130 * b + 32768) >> 16)]
            assert list(got[j, i]) == [min(max(v, 0), 255) for v in exp] + [255]


@pytest.mark.parametrize("w,h,n,kmax,quality", [(64, 48, 10, 0, 75), (50, 34, 9, 3, 40), (33, 17, 6, 1, 90), (16, 16, 1, 0, 75)])
def test_anim_encoder_files_read_by_libwebp(oracle, w, h, n, kmax, quality):
    """AnimEncoder over the oracle codec: sub-frame rectangles, dispose candidates, merged repeats, forced keyframes.  libwebp's own
    demuxer + compositor must show, frame by frame, what our AnimDecoder composites from the oracle-decoded frames, and each
    shown canvas must be close to the source frame (the rectangles cover what changed)."""
    frames = _clip(w, h, n)
    durs = [30 + 10 * i for i in range(n)]
    buf = io.BytesIO()
    enc = animation.AnimEncoder(buf, w, h, animation.EncodeOptions(LoopCount=2, Quality=quality, Kmax=kmax), frame_encoder=_oracle_codec(oracle, quality))
    for f, d in zip(frames, durs):
        enc.AddFrame(f, d)
    enc.Close(simple_encode=_oracle_simple(oracle, quality))
    data = buf.getvalue()
    shown, info = _pillow_canvases(data)
    if n == 1:  # Close writes the still image when it is smaller (animation.go:1190)
        assert data[12:16] == b"VP8 " and len(shown) == 1
        return
    anim = animation.DecodeBytes(data)
    assert (anim.CanvasWidth, anim.CanvasHeight, anim.LoopCount) == (w, h, 2) and info.get("loop") == 2
    merged = sum(1 for i in range(1, n) if np.array_equal(frames[i], frames[i - 1]))
    assert len(anim.Frames) == n - merged == len(shown) and anim.TotalDuration() == sum(durs)
    if kmax == 1:
        assert all(f.OffsetX == 0 and f.OffsetY == 0 and mux.parseVP8Dimensions(f.BitstreamData) == (w, h) for f in anim.Frames)
    else:
        assert any(mux.parseVP8Dimensions(f.BitstreamData) != (w, h) for f in anim.Frames)  # sub-frames were used
    _fancy_frames(oracle, anim)
    dec = animation.AnimDecoder(anim)
    distinct = [f for i, f in enumerate(frames) if i == 0 or not np.array_equal(f, frames[i - 1])]
    for k, (canvas, dur) in enumerate(shown):
        snap, d = dec.NextFrame()
        assert np.array_equal(snap, canvas), k
        assert d == dur
        err = np.abs(snap[..., :3].astype(int) - distinct[k][..., :3].astype(int)).mean()
        assert err < (20 if quality >= 75 else 40), (k, err)  # noise patches at lossy quality; a wrong rectangle would be far off
    assert not dec.HasNext()


def test_anim_encoder_api_behaviour(oracle):
    """animation/animation.go:590-640, 974-1016, 1157-1218: invalid canvas -> no encoder, closed encoder errors, smaller pictures sit
    at (0, 0) of a full canvas, RGB input, AddRawFrame passes bitstreams through, metadata setters; what lies outside the lossy
    path is refused with a message instead of being emulated."""
    codec = _oracle_codec(oracle, 75)
    for w, h in ((0, 10), (10, 0), (16384, 10), (10, 16384)):
        with pytest.raises(animation.AnimError):
            animation.AnimEncoder(io.BytesIO(), w, h)
    for bad in (dict(Lossless=True), dict(AllowMixed=True)):
        with pytest.raises(animation.AnimError, match="outside the GPU lossy path"):
            animation.AnimEncoder(io.BytesIO(), 8, 8, animation.EncodeOptions(**bad), frame_encoder=codec)
    buf = io.BytesIO()
    enc = animation.AnimEncoder(buf, 40, 32, animation.EncodeOptions(Quality=75, LoopCount=70000), frame_encoder=codec)
    assert enc.opts.LoopCount == 0xFFFF  # clampLoopCount
    small = np.full((20, 24, 3), 200, np.uint8)  # RGB, smaller than the canvas
    with pytest.raises(animation.AnimError, match="transparency"):  # the rest of the canvas would be transparent
        enc.AddFrame(small, 40)
    full = np.full((32, 40, 3), 90, np.uint8)
    enc.AddFrame(full, 40)
    moved = full.copy(); moved[4:12, 6:20] = 250
    enc.AddFrame(moved, 60)
    enc.AddFrame(moved, 25)  # identical: merged into the previous frame's duration
    with pytest.raises(animation.AnimError, match="24 bits"):  # the reference would add a transparent 1x1 filler frame here
        enc.AddFrame(moved, 0xFFFFFF)
    raw = mux.riff_payload(oracle.encode(np.dstack([full, np.full((32, 40, 1), 255, np.uint8)])))
    enc.AddRawFrame(raw, 15, 0, 0, animation.BlendNone, animation.DisposeNone)
    enc.SetEXIF(b"Exif\0\0"); enc.SetICCProfile(b"icc"); enc.SetXMP(b"<x/>")
    enc.Close()
    enc.Close()  # a second Close is a no-op (animation.go:1191)
    with pytest.raises(animation.AnimError, match="closed"):
        enc.AddFrame(full, 10)
    a = animation.DecodeBytes(buf.getvalue())
    assert [f.Duration for f in a.Frames] == [40, 85, 15] and (a.ICC, a.EXIF, a.XMP) == (b"icc", b"Exif\0\0", b"<x/>")
    assert a.LoopCount == 0xFFFF and (a.Frames[1].OffsetX, a.Frames[1].OffsetY) == (6, 4)
    assert mux.parseVP8Dimensions(a.Frames[1].BitstreamData) == (14, 8)  # the changed rectangle, even offsets
    shown, info = _pillow_canvases(buf.getvalue())
    assert [d for _, d in shown] == [40, 85, 15] and info.get("icc_profile") == b"icc"


def test_metadata_file_read_by_libwebp(oracle):
    """webp.Encode with ICC / EXIF / XMP (encode.go:955): VP8X container around the unchanged bitstream."""
    img = oracle.synth_image(40, 24, 3)
    plain = oracle.encode(img)
    ext = mux.writeRIFF(mux.FourCCVP8, mux.riff_payload(plain), None, 40, 24, icc=b"profile!", exif=b"Exif\0\0II*\0", xmp=b"<xmp/>")
    shown, info = _pillow_canvases(ext)
    ref, _ = _pillow_canvases(plain)
    assert np.array_equal(shown[0][0], ref[0][0])
    assert info.get("icc_profile") == b"profile!" and info.get("exif") == b"Exif\0\0II*\0" and b"<xmp/>" in info.get("xmp", b"")


@pytest.mark.gpu
def test_gpu_anim_encoder_bytes_and_batching(oracle, gpu_ctx):
    """The GPU codec behind AnimEncoder writes the oracle codec's file byte for byte; the two candidates of a sub-frame are one
    codec call; with Kmax = 1 AddFrames encodes the whole clip in ONE batch and gives the AddFrame-by-AddFrame file."""
    for (w, h, n, kmax, quality) in [(64, 48, 10, 0, 75), (50, 34, 9, 3, 40), (96, 80, 12, 1, 60)]:
        frames = _clip(w, h, n, seed=w)
        durs = [20 + i for i in range(n)]
        out = []
        for codec in (None, _oracle_codec(oracle, quality)):
            buf = io.BytesIO()
            enc = animation.AnimEncoder(buf, w, h, animation.EncodeOptions(Quality=quality, Kmax=kmax), frame_encoder=codec, ctx=gpu_ctx)
            for f, d in zip(frames, durs):
                enc.AddFrame(f, d)
            enc.Close(simple_encode=None if codec is None else _oracle_simple(oracle, quality))
            out.append(buf.getvalue())
        assert out[0] == out[1], (w, h, kmax)
        buf = io.BytesIO()
        enc = animation.AnimEncoder(buf, w, h, animation.EncodeOptions(Quality=quality, Kmax=kmax), ctx=gpu_ctx)
        enc.AddFrames(frames, durs)
        enc.Close()
        assert buf.getvalue() == out[0]
        if kmax == 1:
            assert enc.codec_calls == 1


@pytest.mark.gpu
def test_gpu_animation_decode_frames_batched(oracle, gpu_ctx):
    """Animation.DecodeFrames: every frame of the file in one GPU batch per frame size == the oracle's planes through ycbcrToNRGBA;
    the composited canvases follow."""
    w, h, n = 80, 64, 9
    frames = _clip(w, h, n, seed=9)
    buf = io.BytesIO()
    enc = animation.AnimEncoder(buf, w, h, animation.EncodeOptions(Quality=70, Kmax=4), ctx=gpu_ctx)
    enc.AddFrames(frames, [50] * n)
    enc.Close()
    anim = animation.DecodeBytes(buf.getvalue())
    anim.DecodeFrames(ctx=gpu_ctx)
    for f in anim.Frames:
        fw, fh, y, u, v = oracle.decode(mux.writeRIFFSimple(mux.FourCCVP8, f.BitstreamData))
        assert np.array_equal(f.Image, animation.ycbcrToNRGBA(y[:fh, :fw], u[:(fh + 1) // 2, :(fw + 1) // 2], v[:(fh + 1) // 2, :(fw + 1) // 2]))
    dec = animation.AnimDecoder(anim)
    k = 0
    while dec.HasNext():
        snap, _ = dec.NextFrame()
        assert snap.shape == (h, w, 4) and bool((snap[..., 3] == 255).all())
        k += 1
    assert k == len(anim.Frames)


@pytest.mark.gpu
def test_gpu_encode_with_metadata(oracle, gpu_ctx):
    """webp.Encode with EncoderOptions.ICC / EXIF / XMP: the extended container around the oracle's bitstream; it decodes."""
    img = oracle.synth_image(72, 40, 4)
    o = webp_b200.DefaultOptions()
    o.ICC, o.EXIF, o.XMP = b"icc-bytes", b"Exif\0\0MM\0*", b"<x:xmpmeta/>"
    buf = io.BytesIO()
    webp_b200.Encode(buf, img, o, ctx=gpu_ctx)
    exp = mux.writeRIFFExtended(mux.FourCCVP8, mux.riff_payload(oracle.encode(img)), None, 72, 40, o.ICC, o.EXIF, o.XMP)
    assert buf.getvalue() == exp
    got = webp_b200.Decode(buf.getvalue(), ctx=gpu_ctx)
    _, _, y, u, v = oracle.decode(oracle.encode(img))
    assert np.array_equal(got.Y, y[:40, :72])
