"""GPU parity, operator surface: every batched dsp entry point vs the oracle on the same seeded inputs, bit-exact.
Pattern and input ranges follow internal/dsp/simd_test.go (pixels 0..255, coefficients +-1000 / +-2048, WHT +-256)."""
import ctypes as C

import numpy as np
import pytest

from webp_b200 import dsp

pytestmark = pytest.mark.gpu
N = 20000


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_ftransform(oracle, gpu_ctx):
    rng = np.random.RandomState(1)
    src = rng.randint(0, 256, (N, 16)).astype(np.uint8); ref = rng.randint(0, 256, (N, 16)).astype(np.uint8)
    src[:100] = 255; ref[:100] = 0; src[100:200] = 0; ref[100:200] = 255
    exp = np.zeros((N, 16), np.int16)
    oracle.lib().orc_ftransform_batch(N, _p(src), _p(ref), _p(exp))
    assert np.array_equal(dsp.FTransformBatch(src, ref, gpu_ctx), exp)


@pytest.mark.parametrize("amp", [1000, 2048, 32000])
def test_itransform(oracle, gpu_ctx, amp):
    rng = np.random.RandomState(2)
    ref = rng.randint(0, 256, (N, 16)).astype(np.uint8)
    co = rng.randint(-amp, amp + 1, (N, 16)).astype(np.int16)
    co[: N // 4, 2:] = 0  # AC3-shaped and DC-only blocks
    co[: N // 8, 1:] = 0
    exp = np.zeros((N, 16), np.uint8)
    oracle.lib().orc_itransform_batch(N, _p(ref), _p(co), _p(exp))
    assert np.array_equal(dsp.ITransformBatch(ref, co, gpu_ctx), exp)


def test_wht(oracle, gpu_ctx):
    rng = np.random.RandomState(3)
    a = rng.randint(-2040, 2041, (N, 16)).astype(np.int16)
    exp = np.zeros((N, 16), np.int16)
    oracle.lib().orc_fwht_batch(N, _p(a), _p(exp))
    assert np.array_equal(dsp.FTransformWHTBatch(a, gpu_ctx), exp)
    b = rng.randint(-20000, 20001, (N, 16)).astype(np.int16)
    oracle.lib().orc_iwht_batch(N, _p(b), _p(exp))
    assert np.array_equal(dsp.TransformWHTBatch(b, gpu_ctx), exp)


def test_sse_tdisto(oracle, gpu_ctx):
    rng = np.random.RandomState(4)
    a = rng.randint(0, 256, (N, 16)).astype(np.uint8); b = rng.randint(0, 256, (N, 16)).astype(np.uint8)
    exp = np.zeros(N, np.int32)
    oracle.lib().orc_sse4x4_batch(N, _p(a), _p(b), _p(exp))
    assert np.array_equal(dsp.SSE4x4Batch(a, b, gpu_ctx), exp)
    oracle.lib().orc_tdisto4x4_batch(N, _p(a), _p(b), _p(exp))
    assert np.array_equal(dsp.TDisto4x4Batch(a, b, gpu_ctx), exp)


def test_pred4_all_modes(oracle, gpu_ctx):
    rng = np.random.RandomState(5)
    c = rng.randint(0, 256, (N, 13)).astype(np.uint8)
    c[:50] = 0; c[50:100] = 255
    exp = np.zeros((N, 10, 16), np.uint8)
    oracle.lib().orc_pred4_batch(N, _p(c), _p(exp))
    assert np.array_equal(dsp.PredLuma4Batch(c, gpu_ctx), exp)


@pytest.mark.parametrize("size", [16, 8])
def test_pred_square_all_modes(oracle, gpu_ctx, size):
    rng = np.random.RandomState(8)
    n = 3000
    c = rng.randint(0, 256, (n, 1 + 2 * size)).astype(np.uint8)
    c[:20] = 0; c[20:40] = 255
    exp = np.zeros((n, 7, size, size), np.uint8)
    oracle.lib().orc_pred_square_batch(n, size, _p(c), _p(exp))
    assert np.array_equal(dsp.PredSquareBatch(c, size, gpu_ctx), exp)


@pytest.mark.parametrize("dc_q,ac_q,qtype,sharpen,first", [(24, 30, 0, 1, 0), (24, 30, 0, 1, 1), (48, 46, 1, 0, 0), (21, 27, 2, 0, 0),
                                                             (4, 4, 0, 1, 0), (157, 284, 0, 1, 1)])
def test_quantize(oracle, gpu_ctx, dc_q, ac_q, qtype, sharpen, first):
    rng = np.random.RandomState(6)
    a = rng.randint(-2048, 2049, (N, 16)).astype(np.int16)
    a[: N // 2] = rng.randint(-64, 65, (N // 2, 16))
    exp = np.zeros((N, 16), np.int16); enz = np.zeros(N, np.int32)
    oracle.lib().orc_quantize_batch(N, _p(a), dc_q, ac_q, qtype, sharpen, first, _p(exp), _p(enz))
    got, nz = dsp.QuantizeCoeffsBatch(a, dc_q, ac_q, qtype, sharpen, first, gpu_ctx)
    assert np.array_equal(got, exp) and np.array_equal(nz, enz)


@pytest.mark.parametrize("dc_q,ac_q,first,ctx_type,lam", [(24, 30, 0, 3, 787), (24, 30, 1, 0, 529), (8, 9, 0, 3, 70), (100, 120, 1, 0, 9000)])
def test_trellis_and_token_cost(oracle, gpu_ctx, dc_q, ac_q, first, ctx_type, lam):
    rng = np.random.RandomState(7)
    n = 8000
    a = rng.randint(-2048, 2049, (n, 16)).astype(np.int16)
    a[: n // 2] = (rng.randn(n // 2, 16) * 40).astype(np.int16)
    a[: n // 8] = (rng.randn(n // 8, 16) * 6).astype(np.int16)
    if first:
        a[:, 0] = 0
    c0 = rng.randint(0, 3, n).astype(np.int32)
    exp = np.zeros((n, 16), np.int16); enz = np.zeros(n, np.int32)
    oracle.lib().orc_trellis_batch(n, _p(a), dc_q, ac_q, 0, 1, first, ctx_type, _p(c0), lam, _p(exp), _p(enz))
    got, nz = dsp.TrellisQuantizeBlockBatch(a, dc_q, ac_q, 0, 1, first, ctx_type, c0, lam, gpu_ctx)
    assert np.array_equal(got, exp) and np.array_equal(nz, enz)
    ecost = np.zeros(n, np.int32)
    oracle.lib().orc_token_cost_batch(n, _p(exp), _p(enz), ctx_type, _p(c0), first, _p(ecost))
    assert np.array_equal(dsp.TokenCostForCoeffsBatch(exp, enz, ctx_type, c0, first, gpu_ctx), ecost)


@pytest.mark.parametrize("w,h", [(64, 48), (100, 70), (37, 21), (7, 5), (130, 71), (8, 8), (1536, 1024)])
def test_cleanup_transparent_area(oracle, gpu_ctx, w, h):
    """cleanupTransparentAreaLossy (encode.go:788-890): smoothing of partly transparent 8x8 blocks, flattening of runs of fully
    transparent ones, remainders smoothened only -- a batch of images against the oracle, bit-exact."""
    from test_oracle import alpha_test_image
    imgs = np.stack([alpha_test_image(oracle, w, h, s) for s in (1, 2, 3)])
    imgs[2, ..., 3] = 255  # one fully opaque image: untouched
    out = dsp.CleanupTransparentArea(imgs, gpu_ctx)
    for k in range(3):
        assert np.array_equal(out[k], oracle.cleanup_transparent(imgs[k]))
    assert np.array_equal(out[2], imgs[2])


def test_sse_tdisto_16x16(oracle, gpu_ctx):
    rng = np.random.RandomState(21)
    n = 4000
    a = rng.randint(0, 256, (n, 16, 16)).astype(np.uint8); b = np.clip(a.astype(np.int32) + rng.randint(-40, 41, a.shape), 0, 255).astype(np.uint8)
    b[:50] = 255 - a[:50]
    e1 = np.zeros(n, np.int32); e2 = np.zeros(n, np.int32)
    oracle.lib().orc_sse16x16_batch(n, _p(a), _p(b), _p(e1)); oracle.lib().orc_tdisto16x16_batch(n, _p(a), _p(b), _p(e2))
    assert np.array_equal(dsp.SSE16x16Batch(a, b, gpu_ctx), e1)
    assert np.array_equal(dsp.TDisto16x16Batch(a, b, gpu_ctx), e2)


def test_dequant_and_ftransform2(oracle, gpu_ctx):
    rng = np.random.RandomState(22)
    lv = rng.randint(-2047, 2048, (N, 16)).astype(np.int16)
    for dc_q, ac_q in ((8, 8), (37, 43), (157, 157), (132, 284)):  # the last pair overflows int16 for big levels: truncation as the reference
        exp = np.zeros((N, 16), np.int16)
        oracle.lib().orc_dequant_batch(N, _p(lv), dc_q, ac_q, _p(exp))
        assert np.array_equal(dsp.DequantCoeffsBatch(lv, dc_q, ac_q, gpu_ctx), exp)
    src = rng.randint(0, 256, (N // 2, 2, 16)).astype(np.uint8); ref = rng.randint(0, 256, (N // 2, 2, 16)).astype(np.uint8)
    exp = np.zeros((N // 2, 2, 16), np.int16)
    oracle.lib().orc_ftransform_batch(N, _p(src), _p(ref), _p(exp))  # FTransform2 == FTransform on both blocks (dsp.go:14)
    assert np.array_equal(dsp.FTransform2Batch(src, ref, gpu_ctx), exp)


@pytest.mark.parametrize("kind", ["Transform", "TransformDC", "TransformAC3", "TransformUV", "TransformDCUV"])
def test_decoder_transforms(oracle, gpu_ctx, kind):
    rng = np.random.RandomState(23)
    n = 6000
    uv = kind.endswith("UV")
    co = rng.randint(-2048, 2049, (n, 4, 16) if uv else (n, 16)).astype(np.int16)
    if kind == "TransformAC3":
        m = np.zeros(16, bool); m[[0, 1, 4]] = True
        co[:, ~m] = 0
    if kind in ("TransformDC", "TransformDCUV"):
        co[..., 1:] = 0
    ref = rng.randint(0, 256, (n, 8, 8) if uv else (n, 16)).astype(np.uint8)
    exp = np.zeros_like(ref)
    oracle.lib().orc_dec_transform_batch(n, dsp.DEC_TRANSFORMS[kind], _p(co), _p(ref), _p(exp))
    assert np.array_equal(dsp.DecTransformBatch(kind, co, ref, gpu_ctx), exp)
    if kind in ("TransformDC", "TransformAC3"):  # the short cuts equal the full transform on their input shapes (decode_frame.go:22-44)
        assert np.array_equal(exp, dsp.DecTransformBatch("Transform", co, ref, gpu_ctx))


@pytest.mark.parametrize("kind", dsp.FILTERS)
def test_loop_filter_set(oracle, gpu_ctx, kind):
    rng = np.random.RandomState(24 + dsp.FILTERS.index(kind))
    n = 3000
    base = rng.randint(0, 256, (n, 1, 1)).astype(np.int32)
    tiles = np.clip(base + rng.randint(-30, 31, (n, 24, 24)), 0, 255).astype(np.uint8)  # smooth-ish tiles so that the filters fire
    tiles[: n // 3] = rng.randint(0, 256, (n // 3, 24, 24))
    for thresh, ithresh, hev in ((10, 5, 1), (40, 20, 2), (63, 30, 0), (1, 1, 0)):
        exp = tiles.copy()
        oracle.lib().orc_filter_batch(n, dsp.FILTERS.index(kind), _p(exp), thresh, ithresh, hev)
        got = dsp.FilterBatch(kind, tiles, thresh, ithresh, hev, gpu_ctx)
        assert np.array_equal(got, exp), (kind, thresh)
        if thresh >= 40:
            assert not np.array_equal(got, tiles)  # the case does filter something


@pytest.mark.parametrize("width,channels", [(1, 3), (2, 4), (3, 3), (4, 4), (17, 3), (64, 4), (255, 4)])
def test_upsample_line_pair(oracle, gpu_ctx, width, channels):
    rng = np.random.RandomState(40 + width)
    n, cw = 300, (width + 1) // 2
    ty, by = rng.randint(0, 256, (2, n, width)).astype(np.uint8)
    tu, tv, bu, bv = rng.randint(0, 256, (4, n, cw)).astype(np.uint8)
    at, ab = rng.randint(0, 256, (2, n, width)).astype(np.uint8)
    for with_bot in (True, False):
        for with_alpha in ((False, True) if channels == 4 else (False,)):
            et = np.zeros((n, width, channels), np.uint8); eb = np.zeros_like(et)
            oracle.lib().orc_upsample_line_pair_batch(n, width, _p(ty), _p(by) if with_bot else None, _p(tu), _p(tv), _p(bu), _p(bv),
                                                      _p(at) if with_alpha else None, _p(ab) if with_alpha else None, channels, _p(et), _p(eb))
            gt, gb = dsp.UpsampleLinePairBatch(ty, by if with_bot else None, tu, tv, bu, bv, channels, at if with_alpha else None,
                                               ab if with_alpha else None, gpu_ctx)
            assert np.array_equal(gt, et)
            if with_bot:
                assert np.array_equal(gb, eb)


def test_boolean_coder_real_and_adversarial_streams(oracle, gpu_ctx):
    """wgpu_dsp_boolcode_batch (the encoder's chunk-parallel device coder) vs VP8BitWriter over the same tokens: real token partitions,
    then streams built to break it (probability 0 / 255, carries through 0xff runs, slow-merging range states that need extra
    relaxation rounds, lengths around the chunk size, an empty partition)."""
    rng = np.random.default_rng(11)
    streams = [oracle.encode_tokens(oracle.synth_image(352, 288, 7, kind=2))[0], oracle.encode_tokens(oracle.synth_image(64, 48, 8, kind=1))[0]]
    for n in (0, 1, 9, 4095, 4096, 8191, 8192, 8193, 20000, 100000):
        t = oracle.adversarial_tokens(rng, n, int(rng.integers(0, 6)))
        if n > 9000:
            cut = int(rng.integers(1, n))
            t[cut:] = oracle.adversarial_tokens(rng, n - cut, int(rng.integers(0, 6)))
        streams.append(t)
    streams.append(oracle.adversarial_tokens(rng, 300000, 4))  # near-certain zeros only: the range states barely merge
    streams.append(oracle.adversarial_tokens(rng, 200000, 3))
    got, rounds = dsp.BoolCodeBatch(streams, gpu_ctx)
    for i, (g, t) in enumerate(zip(got, streams)):
        assert np.array_equal(g, oracle.boolcode(t)), "stream %d (%d tokens)" % (i, len(t))
    assert rounds >= 1


def test_boolean_coder_many_partitions(oracle, gpu_ctx):
    rng = np.random.default_rng(12)
    streams = [oracle.adversarial_tokens(rng, int(rng.integers(0, 40000)), int(rng.integers(0, 6))) for _ in range(300)]
    got, _ = dsp.BoolCodeBatch(streams, gpu_ctx)
    for i, (g, t) in enumerate(zip(got, streams)):
        assert np.array_equal(g, oracle.boolcode(t)), "stream %d (%d tokens)" % (i, len(t))

