"""GPU parity, pipeline level: import, analysis, mode search + serialisation (WebP bytes), decoder
reconstruction + loop filter, fancy upsampler, SSE/SSIM -- all through the C ABI, bit-exact against the oracle
(SSIM within 1e-6 relative, BASELINE.json north_star)."""
import ctypes as C
import io
import os

import numpy as np
import pytest

import webp_b200
from webp_b200 import dsp, native

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _opts(**kw):
    o = webp_b200.DefaultOptions()
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _ocfg(oracle, o):
    c = webp_b200.webp.lossy_config(o)
    return oracle.default_cfg(**{f: getattr(c, f) for f, _ in c._fields_})


def _fetch(ctx, i, w, h):
    mbw, mbh = (w + 15) >> 4, (h + 15) >> 4
    nmb = mbw * mbh
    t = dict(mb_hdr=np.zeros((nmb, 8), np.uint8), mb_modes=np.zeros((nmb, 16), np.uint8), mb_nz=np.zeros((nmb, 24), np.uint8),
             mb_coeffs=np.zeros((nmb, 400), np.int16), recon_y=np.zeros((mbh * 16, mbw * 16), np.uint8),
             recon_u=np.zeros((mbh * 8, mbw * 8), np.uint8), recon_v=np.zeros((mbh * 8, mbw * 8), np.uint8),
             src_y=np.zeros((mbh * 16, mbw * 16), np.uint8), src_u=np.zeros((mbh * 8, mbw * 8), np.uint8),
             src_v=np.zeros((mbh * 8, mbw * 8), np.uint8), alphas=np.zeros(nmb, np.uint8))
    ctx.check(native.lib().wgpu_enc_fetch(ctx.handle, i, *[t[k].ctypes.data for k in (
        "mb_hdr", "mb_modes", "mb_nz", "mb_coeffs", "recon_y", "recon_u", "recon_v", "src_y", "src_u", "src_v", "alphas")]))
    return t


def first_diff(name, got, exp):
    if np.array_equal(got, exp):
        return None
    idx = np.argwhere(got != exp)
    return "%s: %d mismatches, first at %s got %s exp %s" % (name, len(idx), idx[0].tolist(), got[tuple(idx[0])], exp[tuple(idx[0])])


@pytest.mark.parametrize("w,h,has_alpha", [(64, 64, 0), (130, 71, 0), (33, 49, 0), (1536, 1024, 0), (96, 80, 1)])
def test_import_rgba(oracle, gpu_ctx, w, h, has_alpha):
    imgs = np.stack([oracle.synth_image(w, h, i) for i in range(3)])
    if has_alpha:
        rng = np.random.RandomState(9)
        imgs[..., 3] = rng.randint(0, 256, imgs.shape[:3])
        imgs[0, :16, :16, 3] = 0
    y, u, v = dsp.ImportRGBA(imgs, bool(has_alpha), gpu_ctx)
    for i in range(3):
        ey, eu, ev = oracle.import_rgba(imgs[i], bool(has_alpha))
        assert np.array_equal(y[i], ey) and np.array_equal(u[i], eu) and np.array_equal(v[i], ev)


@pytest.mark.parametrize("w,h", [(64, 64), (130, 71), (33, 49), (1, 1), (2, 1), (1, 2), (5, 7), (1536, 1024), (4200, 18)])
def test_import_sharp_yuv(oracle, gpu_ctx, w, h):
    """UseSharpYUV planes (sharpyuv.Convert + importYCbCr): sharp_init / sharp_refine / sharp_finish kernels vs the oracle, incl.
    noise and hard-edged images (3-4 refinement passes), odd sizes, and a row of more than 2048 chroma samples (1024-thread CTA)."""
    rng = np.random.RandomState(w * 31 + h)
    imgs = np.stack([oracle.synth_image(w, h, 1), oracle.synth_image(w, h, 5), rng.randint(0, 256, (h, w, 4)).astype(np.uint8),
                     (rng.randint(0, 2, (h, w, 4)) * 255).astype(np.uint8)])
    y, u, v = dsp.ImportRGBA(imgs, False, gpu_ctx, sharp_yuv=True)
    for i in range(len(imgs)):
        ey, eu, ev = oracle.import_rgba(imgs[i], has_alpha=2)
        errs = [e for e in (first_diff("y", y[i], ey), first_diff("u", u[i], eu), first_diff("v", v[i], ev)) if e]
        assert not errs, "image %d: %s" % (i, "; ".join(errs))


@pytest.mark.parametrize("w,h,idxs,kw", [(128, 96, [0, 1, 2], {}), (100, 70, [1, 2], dict(Method=2, Quality=60)), (130, 71, [2, 5], dict(Method=6, Preprocessing=2)),
                                          (64, 48, [0, 1], dict(TargetSize=800))])
def test_encode_with_sharp_yuv(oracle, gpu_ctx, w, h, idxs, kw):
    # EncoderOptions.UseSharpYUV (encode.go:531-535, TestEdge_SharpYUV edge_cases_test.go:482): same bytes as the oracle, and not the standard import's
    o = _opts(UseSharpYUV=True, **kw)
    imgs = np.stack([oracle.synth_image(w, h, i) for i in idxs])
    files = webp_b200.EncodeBatch(imgs, o, gpu_ctx)
    plain = webp_b200.EncodeBatch(imgs, _opts(**kw), gpu_ctx)
    for k in range(len(idxs)):
        assert files[k] == oracle.encode(imgs[k], _ocfg(oracle, o)), "image %d" % idxs[k]
        assert files[k] != plain[k]


def test_import_rgba_dithered(oracle, gpu_ctx):
    # Preprocessing&2: VP8Random rounding terms in the reference's draw order (encode.go:793-809,925-936; dsp/random.go)
    for (w, h, amp) in [(100, 70, 200), (64, 64, 256), (130, 71, 37)]:
        imgs = np.stack([oracle.synth_image(w, h, i) for i in range(2)])
        y, u, v = dsp.ImportRGBA(imgs, False, gpu_ctx, dither_amp=amp)
        for i in range(2):
            ey, eu, ev = oracle.import_rgba(imgs[i], False, dither_amp=amp)
            assert np.array_equal(y[i], ey) and np.array_equal(u[i], eu) and np.array_equal(v[i], ev)


ENC_CASES = [
    (128, 96, [0, 1, 2], {}),
    (100, 70, [1, 2], {}),
    (256, 192, [1], dict(Segments=1)),
    (128, 96, [1, 2], dict(Method=3)),
    (128, 96, [1, 2], dict(Method=6, Quality=40)),
    (128, 96, [1, 2], dict(Quality=95, SNSStrength=0)),
    (128, 96, [2], dict(Quality=5)),
    (128, 96, [1, 2], dict(FilterType=0, FilterSharpness=5, FilterStrength=80, Preprocessing=1)),
    (128, 128, [2], dict(Partitions=2, Quality=90)),
    (128, 128, [1], dict(Partitions=3)),
    (768, 576, [1], {}),
    # Method < 3: statLoop + serial encodeFrame semantics (BASELINE configs[4]: 256x256 q80 method 2)
    (256, 256, [0, 1, 2], dict(Method=2, Quality=80)),
    (256, 256, [4, 5], dict(Method=2, Quality=80, Segments=1)),
    (100, 70, [1, 2], dict(Method=0)),
    (128, 96, [1, 2], dict(Method=1, Quality=50)),
    (64, 32, [1, 2], dict(Method=2)),
    (16, 16, [2], dict(Method=2, Quality=90)),
    (130, 71, [2], dict(Method=2, Partitions=1, Quality=90)),
    (320, 240, [2], dict(Method=2, Pass=3, Quality=60)),
    (128, 96, [1, 2], dict(Preprocessing=2, Quality=60)),  # dithered import (PresetPhoto sets this bit)
    (128, 96, [1], dict(Preprocessing=3, Quality=90, Method=2)),
    # found by tools/fuzz_parity.py: one macroblock column (odd waves are empty); dithering on partial macroblocks (the luma
    # analysis replicates the last real column itself, the dithered padding is not a replica)
    (9, 116, [2, 10, 4], {}),
    (3, 123, [0, 9], dict(Method=1)),
    (52, 118, [7], dict(Quality=0, Method=6, Preprocessing=2, FilterStrength=100, FilterSharpness=7, FilterType=0, Segments=2)),
    (79, 118, [9, 7, 5], dict(Quality=5, Preprocessing=3, FilterStrength=100, FilterSharpness=3, FilterType=0, Segments=3)),
    # Method >= 3 on fewer than 4 macroblock rows: the reference's serial RD path (all-mode I4 search at Method 3,
    # trial mode context kept only when the search completes, chroma DC error diffusion) -- built up to 96 macroblocks
    (64, 48, [0, 1, 2], {}),
    (100, 40, [1, 2], dict(Method=3)),
    (33, 17, [1, 2], dict(Method=6, Quality=90)),
    (512, 48, [1, 2], dict(Method=3, Quality=30)),
    (512, 48, [0, 2], dict(Method=5, Quality=85, Segments=2)),
    (16, 16, [1, 2], {}),
    (160, 33, [2], dict(Quality=98, SNSStrength=100, Partitions=1)),
]


@pytest.mark.parametrize("w,h,idxs,kw", ENC_CASES)
def test_encode_bytes_and_decisions(oracle, gpu_ctx, w, h, idxs, kw):
    o = _opts(**kw)
    imgs = np.stack([oracle.synth_image(w, h, i) for i in idxs])
    files = webp_b200.EncodeBatch(imgs, o, gpu_ctx)
    for k, i in enumerate(idxs):
        exp, t = oracle.encode(imgs[k], _ocfg(oracle, o), taps=True)
        g = _fetch(gpu_ctx, k, w, h)
        errs = [first_diff("alphas", g["alphas"], t["alphas"]),
                first_diff("segment", g["mb_hdr"][:, 3], t["mb_hdr"][:, 3]),
                first_diff("mb_type", g["mb_hdr"][:, 0], t["mb_hdr"][:, 0]),
                first_diff("hdr", g["mb_hdr"][:, :6], t["mb_hdr"][:, :6]),
                first_diff("modes", g["mb_modes"], t["mb_modes"]),
                first_diff("nz", g["mb_nz"], t["mb_nz"]),
                first_diff("coeffs", g["mb_coeffs"], t["mb_coeffs"]),
                first_diff("recon_y", g["recon_y"][:h, :w], t["recon_y"][:h, :w])]
        errs = [e for e in errs if e]
        assert not errs, "image %d: %s" % (i, "; ".join(errs))
        assert files[k] == exp, "image %d: bitstream differs (%d vs %d bytes)" % (i, len(files[k]), len(exp))


@pytest.mark.parametrize("device_coder", ["0", "1"])
@pytest.mark.parametrize("w,h,idxs,kw", [(128, 96, [0, 1, 2], {}), (768, 576, [1, 2], {}), (100, 70, [2], dict(Quality=98, Method=6)),
                                          (64, 48, [0, 1], {}), (320, 240, [0, 1, 2, 4, 5, 7, 8], dict(Quality=20))])
def test_token_partition_coder_routes(oracle, gpu_ctx, monkeypatch, device_coder, w, h, idxs, kw):
    """The token partition is boolean-coded either on the host (code_token_streams) or on the GPU (boolcode_par.cuh:
    VP8BitWriter PutBit / Flush / Finish, internal/bitio/writer_bool.go:58-150); both must give the oracle's bytes."""
    monkeypatch.setenv("WGPU_DEVICE_CODER", device_coder)
    o = _opts(**kw)
    imgs = np.stack([oracle.synth_image(w, h, i) for i in idxs])
    files = webp_b200.EncodeBatch(imgs, o, gpu_ctx)
    for k in range(len(idxs)):
        assert files[k] == oracle.encode(imgs[k], _ocfg(oracle, o))


def test_device_coder_flat_and_noise_images(oracle, gpu_ctx, monkeypatch):
    # all-skip frames (zero tokens: only the closing flush) and dense noise (long 0xff runs / carries are likelier)
    monkeypatch.setenv("WGPU_DEVICE_CODER", "1")
    rng = np.random.RandomState(5)
    flat = np.full((64, 64, 4), 255, np.uint8)
    noise = rng.randint(0, 256, (64, 64, 4)).astype(np.uint8); noise[..., 3] = 255
    imgs = np.stack([flat, noise])
    for q in (75, 100):
        o = _opts(Quality=q)
        files = webp_b200.EncodeBatch(imgs, o, gpu_ctx)
        for k in range(2):
            assert files[k] == oracle.encode(imgs[k], _ocfg(oracle, o))


def test_analyze_search_route_matches_enc_device(oracle, gpu_ctx):
    """The Go-shim route: GPU analysis -> host segmentation (here: taken from the oracle) -> GPU mode
    search with host-supplied SegmentInfo -> per-MB results identical to the oracle's mbInfo."""
    L = native.lib()
    w, h = 160, 128
    img = oracle.synth_image(w, h, 1)
    _, t = oracle.encode(img, taps=True)
    nmb = ((w + 15) >> 4) * ((h + 15) >> 4)
    opt = webp_b200.webp.lossy_config(webp_b200.DefaultOptions())
    gpu_ctx.check(L.wgpu_enc_upload(gpu_ctx.handle, img.ctypes.data, 1, w, h, w * 4, w * h * 4))
    alphas = np.zeros(nmb, np.uint8); uv_sum = np.zeros(1, np.int64)
    gpu_ctx.check(L.wgpu_enc_analyze(gpu_ctx.handle, C.byref(opt), alphas.ctypes.data, uv_sum.ctypes.data))
    assert np.array_equal(alphas, t["alphas"])
    # setSegmentParams (encode_analysis.go:122): UV quantiser deltas from the global chroma alpha (Go integer division)
    guv = int(uv_sum[0]) // nmb
    num = (guv - 64) * 10
    dq = abs(num) // 70 * (1 if num >= 0 else -1)
    dq = abs(dq * 50) // 100 * (1 if dq >= 0 else -1)
    dq_uv_ac, dq_uv_dc = max(-4, min(6, dq)), -2
    segs = (native.Segment * 4)()
    for i in range(4):
        assert L.wgpu_setup_segment(int(t["seg"][i][0]), dq_uv_dc, dq_uv_ac, 4, 50, C.byref(segs[i])) == 0
        assert (segs[i].lambda_i4, segs[i].lambda_i16, segs[i].lambda_uv, segs[i].lambda_mode) == tuple(int(x) for x in t["seg"][i][4:8])
    seg_map = np.ascontiguousarray(t["mb_hdr"][:, 3])
    gpu_ctx.check(L.wgpu_enc_search(gpu_ctx.handle, segs, seg_map.ctypes.data))
    g = _fetch(gpu_ctx, 0, w, h)
    for k in ("mb_modes", "mb_nz", "mb_coeffs"):
        assert np.array_equal(g[k], t[k]), k
    assert np.array_equal(g["mb_hdr"][:, :6], t["mb_hdr"][:, :6])


def test_encode_test_png(oracle, gpu_ctx):
    # BASELINE.json configs[0]: testdata/test.png q75 m4 (768x576 RGBA, all opaque)
    from PIL import Image
    img = np.array(Image.open(os.path.join(DATA, "test.png")).convert("RGBA"))
    buf = io.BytesIO()
    webp_b200.Encode(buf, img, webp_b200.DefaultOptions(), gpu_ctx)
    assert buf.getvalue() == oracle.encode(img)


RC_CASES = [
    (128, 96, [0, 1, 2], dict(TargetPSNR=40.0)),                 # 48 macroblocks: passes at Q, Q-10, Q-10
    (128, 96, [1, 2], dict(TargetSize=2000)),                    # secant search on the trial frame size
    (160, 128, [0, 2], dict(TargetSize=1500, Method=6)),          # 80 macroblocks
    (100, 70, [1, 2], dict(TargetPSNR=35.0, Method=3, Quality=45)),   # all-mode I4 search, quality crossing 50 (max I4 modes 3 -> 2)
    (64, 48, [0, 1, 2], dict(TargetSize=600, QMin=20, QMax=90, Pass=6)),
    (96, 96, [2], dict(TargetSize=4000, Segments=1, Quality=90)),
    # more than 96 macroblocks: mid-stream probability refreshes feed the RD costs (encode_frame.go:35-57)
    (256, 256, [0, 1, 2], dict(TargetPSNR=40.0)),                  # 256 MBs: refreshes before macroblocks 96 and 193, three passes
    (320, 240, [1, 2], dict(TargetSize=9000, Method=6)),           # trial frames carry tokens recorded under mixed tables
    (640, 48, [0, 2], dict(Method=4)),                             # 120 MBs on 3 rows: serial path without rate control
    (1600, 40, [1], dict(Method=3, Quality=60)),                   # 300 MBs, all-mode I4 search
    (200, 150, [0, 1, 2, 4], dict(TargetSize=3000, Quality=50, QMin=10, QMax=80)),
    # Method < 3: statLoop, then the search passes over the non-RD serial path; refresh points see the previous pass below them
    (128, 96, [0, 1, 2], dict(Method=2, TargetSize=1500)),
    (256, 256, [1, 2, 5], dict(Method=2, TargetSize=5000, Quality=80)),
    (320, 240, [0, 2], dict(Method=0, TargetPSNR=40.0, Pass=2)),
    (400, 300, [1, 4], dict(Method=1, TargetSize=12000, Quality=60, Segments=2)),
]


@pytest.mark.parametrize("w,h,idxs,kw", RC_CASES)
def test_rate_control_passes(oracle, gpu_ctx, w, h, idxs, kw):
    """TargetSize / TargetPSNR: the reference's doSearch loop (internal/lossy/encode.go:1338-1374, adjustQuantForTarget
    :1544) over the serial RD encodeFrame, images of one batch converging independently; bytes and per-MB data == oracle."""
    o = _opts(**kw)
    imgs = np.stack([oracle.synth_image(w, h, i) for i in idxs])
    files = webp_b200.EncodeBatch(imgs, o, gpu_ctx)
    for k, i in enumerate(idxs):
        exp, t = oracle.encode(imgs[k], _ocfg(oracle, o), taps=True)
        g = _fetch(gpu_ctx, k, w, h)
        for d in (g, t):  # fields the macroblock type leaves unused keep whatever an earlier pass wrote (in the reference too)
            i4 = d["mb_hdr"][:, 0] == 1
            d["mb_modes"][~i4] = 0
            d["mb_coeffs"][i4, 384:] = 0
            d["mb_hdr"][i4, 5] = 0
            d["mb_hdr"][i4, 1] = 0
        errs = [first_diff("segment", g["mb_hdr"][:, 3], t["mb_hdr"][:, 3]),
                first_diff("hdr", g["mb_hdr"][:, :6], t["mb_hdr"][:, :6]),
                first_diff("modes", g["mb_modes"], t["mb_modes"]),
                first_diff("coeffs", g["mb_coeffs"], t["mb_coeffs"])]
        # a search that has not converged after its last pass restores the source planes (restoreSourcePixels), so the
        # reference's planes hold the source then, not the reconstruction
        if not np.array_equal(t["recon_y"], t["src_y"]):
            errs.append(first_diff("recon_y", g["recon_y"][:h, :w], t["recon_y"][:h, :w]))
        errs = [e for e in errs if e]
        assert not errs, "image %d: %s" % (i, "; ".join(errs))
        assert files[k] == exp, "image %d: bitstream differs (%d vs %d bytes)" % (i, len(files[k]), len(exp))


@pytest.mark.parametrize("device_coder", ["0", "1"])
@pytest.mark.parametrize("w,h,idxs,kw", [
    (81, 47, [10, 11], dict(Quality=20, TargetSize=8000, Preprocessing=3, SNSStrength=0, FilterStrength=60, FilterSharpness=3, FilterType=0, Pass=1)),
    (94, 155, [6, 8], dict(Quality=49, Method=5, TargetSize=300, Preprocessing=1, SNSStrength=100, FilterStrength=100, Segments=4)),
    (100, 70, [5, 7, 2], dict(Quality=20, Method=5, TargetSize=8000, Preprocessing=2, SNSStrength=0, FilterStrength=60, FilterType=0, Segments=4, Pass=1))])
def test_rate_control_unconverged_last_pass_header(oracle, gpu_ctx, monkeypatch, device_coder, w, h, idxs, kw):
    """A search that has not converged after its last pass re-derives the segment parameters, and the frame header (partition 0)
    is written from THAT state while the macroblocks keep the last pass's data.  Found by tools/fuzz_parity.py (seed 7) when
    partition 0 moved to the device: its segment map has to be the host's final one.  Both coder routes, bytes == oracle."""
    monkeypatch.setenv("WGPU_DEVICE_CODER", device_coder)
    o = _opts(**kw)
    imgs = np.stack([oracle.synth_image(w, h, i) for i in idxs])
    files = webp_b200.EncodeBatch(imgs, o, gpu_ctx)
    for k in range(len(idxs)):
        assert files[k] == oracle.encode(imgs[k], _ocfg(oracle, o)), "image %d" % idxs[k]


def test_config4_4k_target_psnr_method6(oracle, gpu_ctx):
    """BASELINE configs[3]: 3840x2160, Method 6, TargetPSNR multi-pass -- three serial RD passes (Q, Q-10, Q-10: the
    reference's PSNR reading is always 99 dB) of 32 400 macroblocks with seven probability refreshes each."""
    w, h = 3840, 2160
    img = oracle.synth_image(w, h, 1)
    o = _opts(Method=6, TargetPSNR=42.0)
    files = webp_b200.EncodeBatch(img[None], o, gpu_ctx)
    assert files[0] == oracle.encode(img, _ocfg(oracle, o))


def test_encode_rejections(gpu_ctx):
    img = np.full((64, 64, 4), 255, np.uint8)
    wide = np.full((40, 16 * 97, 4), 255, np.uint8)
    with pytest.raises(native.WebPGPUError) as e:  # serial RD path with refreshes: one token partition only
        webp_b200.EncodeBatch(wide[None], _opts(Partitions=2), gpu_ctx)
    assert e.value.code == native.ERR_UNSUPPORTED
    big = np.full((256, 256, 4), 255, np.uint8)
    with pytest.raises(native.WebPGPUError) as e:  # the refresh schedule is built for one token partition only
        webp_b200.EncodeBatch(big[None], _opts(TargetPSNR=40.0, Partitions=1), gpu_ctx)
    assert e.value.code == native.ERR_UNSUPPORTED
    with pytest.raises(native.WebPGPUError) as e:  # rate control x token partitions: stale per-macroblock token starts across passes
        webp_b200.EncodeBatch(img[None], _opts(TargetPSNR=42.0, Partitions=2), gpu_ctx)
    assert e.value.code == native.ERR_UNSUPPORTED
    with pytest.raises(webp_b200.WebPError):
        webp_b200.EncodeBatch(img[None], _opts(Lossless=True), gpu_ctx)


DEC_CASES = [(128, 96, (2, 1, 0), {}), (100, 70, (2, 1, 0), {}), (130, 71, (2, 1, 0), dict(segments=1)),
             (128, 96, (2, 1, 0), dict(filter_type=0, segments=1)), (128, 96, (2, 1, 0), dict(filter_strength=0)),
             (160, 112, (2, 1, 0), dict(filter_sharpness=6, segments=1, filter_strength=100)),
             # multi-partition streams of images with skipped macroblocks are corrupt in the reference itself
             # (encode_token.go:90 marks MB starts only for non-skipped MBs, DESIGN.md "reference quirks")
             (128, 128, (2, 5, 8), dict(partitions=2, quality=90)), (128, 128, (2,), dict(partitions=3, quality=95)),
             (768, 576, (2, 1, 0), dict(segments=1)), (64, 64, (2, 1, 0), dict(quality=10))]


@pytest.fixture(params=["0", "1"], ids=["host_parser", "gpu_parser"])
def parser_route(request, monkeypatch):
    """Macroblock parsing (intra modes + coefficient tokens) on the host (host_dec.h::parse_frame) or on the GPU
    (dec_parse_kernel: decode_tree.go:35, decode_mb.go:111-313); the library picks by batch size, the tests force both."""
    monkeypatch.setenv("WGPU_DEVICE_PARSER", request.param)
    return request.param


@pytest.mark.parametrize("w,h,idxs,kw", DEC_CASES)
def test_decode_planes_and_nrgba(oracle, gpu_ctx, parser_route, w, h, idxs, kw):
    streams = [oracle.encode(oracle.synth_image(w, h, i), oracle.default_cfg(**kw)) for i in idxs]
    gw, gh, y, u, v, rgba = webp_b200.webp.decode_padded(streams, nrgba=True, ctx=gpu_ctx)
    assert (gw, gh) == (w, h)
    for i, s in enumerate(streams):
        _, _, ey, eu, ev = oracle.decode(s)
        errs = [first_diff("y", y[i], ey), first_diff("u", u[i], eu), first_diff("v", v[i], ev),
                first_diff("nrgba", rgba[i], oracle.build_nrgba(w, h, ey, eu, ev))]
        errs = [e for e in errs if e]
        assert not errs, "stream %d: %s" % (i, "; ".join(errs))


def test_decode_reference_fixtures_and_foreign_stream(oracle, gpu_ctx, parser_route):
    for name in ("blue_16x16_lossy.webp", "red_4x4_lossy.webp"):
        data = open(os.path.join(DATA, name), "rb").read()
        w, h, y, u, v, rgba = webp_b200.webp.decode_padded([data], nrgba=True, ctx=gpu_ctx)
        _, _, ey, eu, ev = oracle.decode(data)
        assert np.array_equal(y[0], ey) and np.array_equal(u[0], eu) and np.array_equal(v[0], ev)
        assert np.array_equal(rgba[0], oracle.build_nrgba(w, h, ey, eu, ev))
    img = webp_b200.Decode(io.BytesIO(open(os.path.join(DATA, "blue_16x16_lossy.webp"), "rb").read()), gpu_ctx)
    assert img.Y.shape == (16, 16) and img.Cb.shape == (8, 8)
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(oracle.synth_image(200, 120, 1)[..., :3]).save(buf, "WEBP", quality=60, method=4)
    data = buf.getvalue()
    w, h, y, u, v, _ = webp_b200.webp.decode_padded([data], ctx=gpu_ctx)
    _, _, ey, eu, ev = oracle.decode(data)
    assert np.array_equal(y[0], ey) and np.array_equal(u[0], eu) and np.array_equal(v[0], ev)


def test_decode_errors(gpu_ctx, oracle, parser_route):
    data = oracle.encode(oracle.synth_image(64, 64, 1))
    with pytest.raises(webp_b200.WebPError):  # RIFF chunk longer than the file
        webp_b200.webp.decode_padded([data[:len(data) // 2]], ctx=gpu_ctx)
    with pytest.raises(native.WebPGPUError) as e:  # raw VP8 frame cut in the token partition
        webp_b200.webp.decode_padded([data[20:20 + (len(data) - 20) // 2]], ctx=gpu_ctx)
    assert e.value.code == native.ERR_BITSTREAM
    with pytest.raises(RuntimeError):
        oracle.decode(data[20:20 + (len(data) - 20) // 2])
    other = oracle.encode(oracle.synth_image(80, 64, 1))
    with pytest.raises(native.WebPGPUError):
        webp_b200.webp.decode_padded([data, other], ctx=gpu_ctx)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 2), (5, 3), (16, 16), (33, 49), (130, 71), (1536, 1024)])
def test_upsample_nrgba(oracle, gpu_ctx, w, h):
    rng = np.random.RandomState(w * 31 + h)
    n = 2
    cw, ch = (w + 1) // 2, (h + 1) // 2
    y = rng.randint(0, 256, (n, h, w)).astype(np.uint8)
    u = rng.randint(0, 256, (n, ch, cw)).astype(np.uint8)
    v = rng.randint(0, 256, (n, ch, cw)).astype(np.uint8)
    a = rng.randint(0, 256, (n, h, w)).astype(np.uint8)
    got = dsp.UpsampleNRGBA(y, u, v, w, h, None, gpu_ctx)
    got_a = dsp.UpsampleNRGBA(y, u, v, w, h, a, gpu_ctx)
    for i in range(n):
        assert np.array_equal(got[i], oracle.build_nrgba(w, h, y[i], u[i], v[i]))
        assert np.array_equal(got_a[i], oracle.build_nrgba(w, h, y[i], u[i], v[i], a[i]))


@pytest.mark.parametrize("w,h", [(1, 1), (7, 7), (8, 3), (40, 56), (33, 57), (130, 71), (257, 119), (768, 512), (1536, 1024)])
def test_plane_metrics(oracle, gpu_ctx, w, h):
    rng = np.random.RandomState(w + h)
    a = rng.randint(0, 256, (2, h, w)).astype(np.uint8)
    b = np.clip(a.astype(np.int32) + rng.randint(-12, 13, a.shape), 0, 255).astype(np.uint8)
    b[1] = a[1]
    sse, ssim = dsp.PlaneMetrics(a, b, gpu_ctx)
    for i in range(2):
        assert int(sse[i]) == oracle.plane_sse(a[i], b[i])
        exp = oracle.plane_ssim(a[i], b[i])
        assert abs(ssim[i] - exp) <= 1e-6 * abs(exp)  # north_star tolerance: 1e-6 relative
        assert dsp.PSNRFromSSE(sse[i], w * h) == oracle.lib().orc_psnr_from_sse(int(sse[i]), w * h)


def test_partition_starting_with_0xff_goes_to_the_host_decoder(oracle, gpu_ctx, monkeypatch):
    """A (corrupt) partition whose first byte is 0xff breaks value >> bits <= range, which the GPU parser's narrow window relies
    on (tests/test_decoder_model.py): the library parses such a batch on the host whatever route is asked for, so both routes
    must make the same thing of it (the same planes, or the same refusal)."""
    good = oracle.encode(oracle.synth_image(96, 80, 1))
    part0_len = (good[20] | (good[21] << 8) | (good[22] << 16)) >> 5
    for pos in (20 + 10 + part0_len, 20 + 10):  # first byte of the token partition / of partition 0
        bad = bytearray(good)
        bad[pos] = 0xff
        bad = bytes(bad)
        outcome = []
        for route in ("0", "1"):
            monkeypatch.setenv("WGPU_DEVICE_PARSER", route)
            try:
                _, _, y, u, v, _ = webp_b200.webp.decode_padded([good, bad], ctx=gpu_ctx)
                outcome.append((y[1].tobytes(), u[1].tobytes(), v[1].tobytes()))
            except (native.WebPGPUError, webp_b200.WebPError) as e:
                outcome.append(("error", type(e).__name__))
        assert outcome[0] == outcome[1]


def test_large_batch_takes_gpu_coder_and_parser(oracle, gpu_ctx, monkeypatch):
    """40 images in one call: the library's own choice of routes (>= 32 images: token partitions coded and macroblocks parsed
    on the GPU; 40 = one full warp of partitions + a partial one), bytes and decoded planes against the oracle."""
    monkeypatch.delenv("WGPU_DEVICE_CODER", raising=False)
    monkeypatch.delenv("WGPU_DEVICE_PARSER", raising=False)
    w, h = 96, 80
    imgs = np.stack([oracle.synth_image(w, h, i % 12) for i in range(40)])
    imgs[5] = 255  # a flat image: all macroblocks skipped, an empty token partition
    for kw in ({}, dict(Method=2, Quality=60)):
        o = _opts(**kw)
        files = webp_b200.EncodeBatch(imgs, o, gpu_ctx)
        exp = {}
        for k in range(40):
            key = imgs[k].tobytes()
            if key not in exp:
                exp[key] = oracle.encode(imgs[k], _ocfg(oracle, o))
            assert files[k] == exp[key], "image %d" % k
        _, _, y, u, v, rgba = webp_b200.webp.decode_padded(files, nrgba=True, ctx=gpu_ctx)
        for k in (0, 5, 31, 32, 39):
            _, _, ey, eu, ev = oracle.decode(files[k])
            assert np.array_equal(y[k], ey) and np.array_equal(u[k], eu) and np.array_equal(v[k], ev)
            assert np.array_equal(rgba[k], oracle.build_nrgba(w, h, ey, eu, ev))


def test_round_trip_at_full_size(oracle, gpu_ctx, parser_route):
    """BASELINE configs[1]/[2] shape: 1536x1024 q75 m4 -> own decoder -> libwebp-identical planes; size-independent
    properties: GPU decode of GPU bytes == encoder's own reconstruction (filter off), PSNR >= 30 dB."""
    imgs = np.stack([oracle.synth_image(1536, 1024, i) for i in (0, 1)])
    o = _opts(FilterStrength=0)
    files = webp_b200.EncodeBatch(imgs, o, gpu_ctx)
    g = [_fetch(gpu_ctx, k, 1536, 1024) for k in range(2)]
    _, _, y, u, v, _ = webp_b200.webp.decode_padded(files, ctx=gpu_ctx)
    for k in range(2):
        assert np.array_equal(y[k], g[k]["recon_y"]) and np.array_equal(u[k], g[k]["recon_u"]) and np.array_equal(v[k], g[k]["recon_v"])
        sse, _ = dsp.PlaneMetrics(g[k]["src_y"][None], y[k][None], gpu_ctx)
        assert dsp.PSNRFromSSE(sse[0], 1536 * 1024) >= 30.0
    assert files[1] == oracle.encode(imgs[1], _ocfg(oracle, o))


def test_two_contexts_concurrently_at_bench_size(oracle, monkeypatch):
    """What bench.py times, checked: 64 images 1536x1024 with DEFAULT options (FilterStrength 60, 4 segments, Method 4) through
    TWO contexts running at the same time on one GPU, token partitions coded on the GPU (the coder kernels of one context beside
    the other context's mode-search waves); then the streams decoded concurrently on both contexts, one per macroblock-parser
    route, with the complex loop filter on.  Files, planes and NRGBA against the oracle."""
    import concurrent.futures as cf
    import threading
    monkeypatch.setenv("WGPU_DEVICE_CODER", "1")
    w, h, n = 1536, 1024, 64
    distinct = [np.stack([oracle.synth_image(w, h, 3 * k + j) for j in range(3)]) for k in (0, 1)]  # six different pictures, all classes
    batches = [np.concatenate([d] * ((n + 2) // 3))[:n] for d in distinct]
    ctxs = [native.Context(0), native.Context(0)]
    try:
        o = webp_b200.DefaultOptions()
        files = [None, None]
        go = threading.Barrier(2)

        def enc(i):
            go.wait()
            for _ in range(2):  # twice: the second batch of a context overlaps the other context's coder
                files[i] = webp_b200.EncodeBatch(batches[i], o, ctxs[i])
        ths = [threading.Thread(target=enc, args=(i,)) for i in range(2)]
        [t.start() for t in ths]; [t.join() for t in ths]
        with cf.ThreadPoolExecutor(max_workers=6) as ex:
            exp = list(ex.map(lambda im: oracle.encode(im, _ocfg(oracle, o)), [distinct[i][j] for i in range(2) for j in range(3)]))
        for i in range(2):
            for k in range(n):
                assert files[i][k] == exp[3 * i + k % 3], "context %d image %d" % (i, k)
        # decode: context 0 parses on the host, context 1 on the GPU
        out = [None, None]

        # the parser route is an environment knob read at call time: run the two routes one after the other with the OTHER context
        # busy encoding, so that each decode still shares the GPU with foreign kernels
        for i in range(2):
            os.environ["WGPU_DEVICE_PARSER"] = str(i)
            t = threading.Thread(target=lambda j=1 - i: webp_b200.EncodeBatch(batches[j][:16], o, ctxs[j]))
            t.start()
            out[i] = webp_b200.webp.decode_padded(files[i], nrgba=True, ctx=ctxs[i])
            t.join()
        os.environ.pop("WGPU_DEVICE_PARSER", None)
        for i in range(2):
            _, _, y, u, v, rgba = out[i]
            for k in (0, 1, 2, n - 1):
                _, _, ey, eu, ev = oracle.decode(files[i][k])
                assert np.array_equal(y[k], ey) and np.array_equal(u[k], eu) and np.array_equal(v[k], ev), "planes ctx %d image %d" % (i, k)
                assert np.array_equal(rgba[k], oracle.build_nrgba(w, h, ey, eu, ev)), "nrgba ctx %d image %d" % (i, k)
    finally:
        os.environ.pop("WGPU_DEVICE_PARSER", None)
        for c in ctxs:
            c.close()


def test_alph_and_vp8x_alpha_files_are_rejected(oracle, gpu_ctx):
    """A lossy-with-alpha file (VP8X alpha flag + ALPH chunk) must not decode as if it were opaque (webp.go:323-350 returns NRGBA
    with the real alpha there): the GPU path rejects it with its own message instead of silently dropping the alpha plane."""
    import struct
    vp8 = oracle.encode(oracle.synth_image(32, 32, 1))[20:]  # the VP8 payload of a simple RIFF file
    def chunk(tag, payload):
        return tag + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"")
    vp8x = chunk(b"VP8X", bytes([0x10, 0, 0, 0, 31, 0, 0, 31, 0, 0]))
    body = b"WEBP" + vp8x + chunk(b"ALPH", b"\0" + bytes(32 * 32)) + chunk(b"VP8 ", vp8)
    data = b"RIFF" + struct.pack("<I", len(body)) + body
    with pytest.raises(webp_b200.WebPError) as e:
        webp_b200.DecodeBatch([data], ctx=gpu_ctx)
    assert "ALPH" in str(e.value) or "alpha" in str(e.value)
    with pytest.raises(webp_b200.WebPError):
        webp_b200.DecodeConfig(data)
    # without the alpha flag and chunk the same payload in an extended container still decodes
    body = b"WEBP" + chunk(b"VP8X", bytes([0, 0, 0, 0, 31, 0, 0, 31, 0, 0])) + chunk(b"VP8 ", vp8)
    ok = b"RIFF" + struct.pack("<I", len(body)) + body
    assert webp_b200.DecodeConfig(ok).Width == 32


def test_dec_reconstruct_from_preparsed_macroblocks(oracle, gpu_ctx):
    """the Go decoder keeps its own parser and hands frame-sized MBData + FInfo arrays to wgpu_dec_reconstruct.
    The records come from the product's host parser run on the CPU (hostcheck_parse: coefficients + the 32-byte side record per
    macroblock, checked against the oracle decoder's taps in tests/test_oracle.py); planes and NRGBA must equal the oracle's."""
    H = C.CDLL(os.path.join(os.path.dirname(DATA), "..", "oracle", "_build", "libhostcheck.so"))
    L = native.lib()
    for (w, h, idxs, kw) in [(128, 96, (0, 1, 2), {}), (130, 71, (1, 2), dict(filter_type=0, filter_strength=40)), (320, 240, (2,), dict(filter_strength=0)),
                             (768, 576, (1, 2), dict(quality=40, filter_sharpness=5))]:
        files = [oracle.encode(oracle.synth_image(w, h, i), oracle.default_cfg(**kw)) for i in idxs]
        n, mbw, mbh = len(files), (w + 15) >> 4, (h + 15) >> 4
        nmb = mbw * mbh
        rec = np.zeros((n, nmb, 800), np.uint8)
        ftype = np.zeros(n, np.uint8)
        for i, f in enumerate(files):
            co = np.zeros((nmb, 384), np.int16); me = np.zeros((nmb, 32), np.uint8); dims = (C.c_int * 5)()
            assert H.hostcheck_parse(f, C.c_long(len(f)), co.ctypes.data_as(C.c_void_p), me.ctypes.data_as(C.c_void_p), C.c_long(nmb), dims) == 0
            assert (dims[0], dims[1]) == (w, h)
            rec[i, :, :768] = co.view(np.uint8).reshape(nmb, 768)
            rec[i, :, 768:] = me
            ftype[i] = dims[4]
        gpu_ctx.check(L.wgpu_dec_reconstruct(gpu_ctx.handle, n, w, h, rec.ctypes.data, ftype.ctypes.data, 1))
        y = np.empty((n, mbh * 16, mbw * 16), np.uint8); u = np.empty((n, mbh * 8, mbw * 8), np.uint8); v = np.empty_like(u)
        rgba = np.empty((n, h, w, 4), np.uint8)
        gpu_ctx.check(L.wgpu_dec_fetch(gpu_ctx.handle, y.ctypes.data, u.ctypes.data, v.ctypes.data, y[0].nbytes, u[0].nbytes, rgba.ctypes.data, w * h * 4))
        for i, f in enumerate(files):
            _, _, ey, eu, ev = oracle.decode(f)
            assert np.array_equal(y[i], ey) and np.array_equal(u[i], eu) and np.array_equal(v[i], ev), (w, h, i)
            assert np.array_equal(rgba[i], oracle.build_nrgba(w, h, ey, eu, ev))
    bad = rec.copy(); bad[0, 0, 768 + 8 + 16 + 1] = 9  # uv_mode out of range
    assert L.wgpu_dec_reconstruct(gpu_ctx.handle, n, w, h, bad.ctypes.data, ftype.ctypes.data, 0) == native.ERR_INVALID


def test_enc_stats_for_a_host_side_refresh(oracle, gpu_ctx):
    """wgpu_enc_stats (collectAllStats with the not-yet-encoded macroblocks in their zero state, encode_proba.go:171,
    encode_frame.go:35-57): with no macroblock encoded every one is I16, not skipped, without coefficients -> one EOB event
    per block in context 0 (WHT block: type 1 band 0; 16 luma AC blocks: type 0 band 1; 8 chroma blocks: type 2 band 0); the
    counts grow monotonically with the cut and the full-frame call is repeatable."""
    w, h = 160, 112
    imgs = np.stack([oracle.synth_image(w, h, i) for i in (1, 2)])
    webp_b200.EncodeBatch(imgs, webp_b200.DefaultOptions(), gpu_ctx)
    nmb = 10 * 7
    L = native.lib()
    def stats(cut):
        st = np.zeros((2, 4, 8, 3, 11, 2), np.uint32)
        gpu_ctx.check(L.wgpu_enc_stats(gpu_ctx.handle, cut, st.ctypes.data))
        return st
    z = stats(0)
    exp = np.zeros_like(z)
    exp[:, 1, 0, 0, 0, 0] = nmb; exp[:, 0, 1, 0, 0, 0] = 16 * nmb; exp[:, 2, 0, 0, 0, 0] = 8 * nmb
    assert np.array_equal(z, exp)
    full = stats(nmb)
    assert np.array_equal(full, stats(nmb)) and not np.array_equal(full, z)
    half = stats(nmb // 2)
    assert int(half.sum()) >= int(z.sum()) and int(full.sum()) >= int(half.sum())
    assert L.wgpu_enc_stats(gpu_ctx.handle, nmb + 1, z.ctypes.data) == native.ERR_INVALID


def test_context_on_second_device(oracle):
    """ADVICE: contexts of one process on different GPUs (function attributes and tables are per device).  Needs two GPUs."""
    import ctypes as C
    from webp_b200 import native
    count = C.c_int(0)
    cudart = None
    for name in ("libcudart.so", "libcudart.so.12"):
        try:
            cudart = C.CDLL(name); break
        except OSError:
            continue
    if cudart is None or cudart.cudaGetDeviceCount(C.byref(count)) != 0 or count.value < 2:
        pytest.skip("one GPU visible")
    c0, c1 = native.Context(0), native.Context(1)
    try:
        imgs = np.stack([oracle.synth_image(320, 240, i) for i in (0, 1, 2)])
        o = _opts()
        f1 = webp_b200.EncodeBatch(imgs, o, c1)
        f0 = webp_b200.EncodeBatch(imgs, o, c0)
        o2 = _opts(TargetSize=4000)
        r1 = webp_b200.EncodeBatch(imgs, o2, c1)
        for k in range(3):
            assert f1[k] == f0[k] == oracle.encode(imgs[k], _ocfg(oracle, o))
            assert r1[k] == oracle.encode(imgs[k], _ocfg(oracle, o2))
        w, h, y, u, v, rgba = webp_b200.webp.decode_padded(f1, nrgba=True, ctx=c1)
        for k in range(3):
            _, _, ey, eu, ev = oracle.decode(f1[k])
            assert np.array_equal(y[k], ey) and np.array_equal(rgba[k], oracle.build_nrgba(w, h, ey, eu, ev))
    finally:
        c0.close(); c1.close()


def test_allocator_reuse_across_shapes_and_trim(oracle):
    """Device image-batch allocator (include/webpgpu.h wgpu_ctx_mem_info / wgpu_ctx_trim; internal/pool/pool.go:14-71,
    resetForReuse internal/lossy/encode.go:391): one context is driven through shapes, options and routes that leave every
    kind of stale content behind (larger batch, other dimensions, rate control, Method 2, decode), and each result must be
    what a fresh context gives = the oracle's bytes / planes.  Buffers only grow, in bucket sizes; trim gives the working set
    back, keeps the constant tables and the next call starts from nothing."""
    ctx = native.Context(0)
    L = native.lib()
    assert [L.wgpu_pool_bucket(n) for n in (1, 256, 257, 5000, 1 << 20, (1 << 20) + 1, 3 << 20, (5 << 20) + 7)] == [
        256, 256, 1024, 16384, 1 << 20, (1 << 20) + (1 << 17), 3 << 20, (5 << 20) + (1 << 19)]
    d0, h0, k0 = ctx.mem_info()
    assert d0 > 0 and h0 == 0  # constant tables only
    plan = [(4, 160, 112, {}), (1, 48, 48, {}), (2, 96, 144, dict(TargetSize=1500)), (3, 64, 64, dict(Method=2, Quality=80)),
            (6, 160, 112, dict(Quality=30)), (1, 16, 16, {}), (2, 200, 40, dict(FilterStrength=0, Segments=1))]
    grown = []
    for n, w, h, kw in plan:
        imgs = np.stack([oracle.synth_image(w, h, i, kind=i % 3) for i in range(n)])
        o = _opts(**kw)
        files = webp_b200.EncodeBatch(imgs, o, ctx)
        exp = [oracle.encode(imgs[i], _ocfg(oracle, o)) for i in range(n)]
        assert files == exp, (n, w, h, kw)
        ys, rgba = webp_b200.DecodeBatch(files, nrgba=True, ctx=ctx)
        for i in range(n):
            _, _, y, u, v = oracle.decode(exp[i])
            assert np.array_equal(ys[i].Y, y[:h, :w]) and np.array_equal(rgba[i], oracle.build_nrgba(w, h, y, u, v))
        d, hp, k = ctx.mem_info()
        grown.append(d)
        assert d >= (grown[-2] if len(grown) > 1 else d0)  # grow-only
    ctx.trim()
    assert ctx.mem_info()[0] <= d0 + (1 << 20) and ctx.mem_info()[1] == 0  # tables (+ lazily built ones) stay, working set gone
    n, w, h, kw = plan[0]
    imgs = np.stack([oracle.synth_image(w, h, i, kind=i % 3) for i in range(n)])
    assert webp_b200.EncodeBatch(imgs, _opts(), ctx) == [oracle.encode(imgs[i]) for i in range(n)]
    files = [oracle.encode(imgs[i]) for i in range(n)]
    ctx.trim()  # decode first on a trimmed context
    ys = webp_b200.DecodeBatch(files, ctx=ctx)
    assert all(np.array_equal(ys[i].Y, oracle.decode(files[i])[2][:h, :w]) for i in range(n))
    ctx.close()


def test_wave_order_does_not_change_the_bytes(oracle, gpu_ctx, monkeypatch):
    """The mode search sorts every wave's task list by the pictures' analysis alpha sums, busiest first (EncKernelParams::img_order,
    DESIGN.md 4a): a scheduling choice only.  A batch that mixes the three content classes must give the oracle's bytes in the
    default order, in image order (WGPU_WAVE_ORDER=0) and in the reversed order (-1)."""
    w, h, n = 208, 144, 13
    imgs = np.stack([oracle.synth_image(w, h, i, kind=(i * 5) % 3) for i in range(n)])
    exp = [oracle.encode(imgs[i]) for i in range(n)]
    for order in (None, "0", "-1"):
        if order is None:
            monkeypatch.delenv("WGPU_WAVE_ORDER", raising=False)
        else:
            monkeypatch.setenv("WGPU_WAVE_ORDER", order)
        assert webp_b200.EncodeBatch(imgs, webp_b200.DefaultOptions(), gpu_ctx) == exp, order
