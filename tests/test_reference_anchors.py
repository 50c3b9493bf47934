"""The analytic vectors the reference's own tests hold for the hot path (tests/golden/reference_anchors.json, transcribed from
internal/dsp/upsample_test.go, internal/lossy/encode_test.go, internal/dsp/random_test.go): the oracle on the CPU, the CUDA
batch twins on the GPU."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
A = json.load(open(os.path.join(HERE, "golden", "reference_anchors.json")))


def _yuv_to_rgb(y, u, v):
    """VP8YUVToR/G/B (internal/dsp/yuv.go:71-104), as upsample_test.go's verifyPixelUV computes its expectation."""
    def clip(x):
        return 0 if x < 0 else (255 if x > 16383 else x >> 6)
    mh = lambda a, c: (a * c) >> 8
    return [clip(mh(y, 19077) + mh(v, 26149) - 14234), clip(mh(y, 19077) - mh(u, 6419) - mh(v, 13320) + 8708), clip(mh(y, 19077) + mh(u, 33050) - 17685)]


def _diamond_inputs():
    d = A["upsample_diamond"]
    f = lambda k: np.array([d[k]], np.uint8)
    return d, f("top_y"), f("bot_y"), f("top_u"), f("top_v"), f("bot_u"), f("bot_v")


def _check_diamond(top, bot, d):
    for row, name in ((top, "top"), (bot, "bot")):
        for x, uv in d["expected_uv"][name].items():
            assert row[0, int(x), :3].tolist() == _yuv_to_rgb(128, uv, uv), (name, x)


def _sq(oracle_or_none, quant, bias_ac, bias_dc):
    return dict(dc_q=quant, ac_q=quant)


def test_oracle_upsample_diamond(oracle):
    d, ty, by, tu, tv, bu, bv = _diamond_inputs()
    L = oracle.lib()
    for ch in (3, 4):
        td = np.zeros((1, d["width"], ch), np.uint8); bd = np.zeros_like(td)
        L.orc_upsample_line_pair_batch(1, d["width"], *[a.ctypes.data_as(C.c_void_p) for a in (ty, by, tu, tv, bu, bv)], None, None, ch,
                                       td.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p))
        _check_diamond(td, bd, d)


def test_oracle_quantize_dequant_psnr_random(oracle):
    L = oracle.lib()
    # the reference's examples use a flat quantiser 10 with the Y1 biases (96 DC / 110 AC): type 0, quantiser index chosen so that
    # kDcTable / kAcTable give 10 is not needed -- orc_quantize_batch takes the quantiser values directly
    for q in A["quantize"]:
        vin = np.array([q["in"]], np.int16); out = np.zeros((1, 16), np.int16); nz = np.zeros(1, np.int32)
        L.orc_quantize_batch(1, vin.ctypes.data_as(C.c_void_p), q["quant"], q["quant"], 0, 0, q["first"], out.ctypes.data_as(C.c_void_p), nz.ctypes.data_as(C.c_void_p))
        _check_quant(q, vin, out, nz)
    d = A["dequant"]
    vin = np.array([d["in"]], np.int16); out = np.zeros((1, 16), np.int16)
    L.orc_dequant_batch(1, vin.ctypes.data_as(C.c_void_p), d["dc_quant"], d["quant"], out.ctypes.data_as(C.c_void_p))
    assert out[0, 0] == d["expect"]["out0"] and out[0, 1] == d["expect"]["out1"]
    for c in A["psnr"]["cases"]:
        assert c["min"] <= oracle.lib().orc_psnr_from_sse(c["mse"], c["size"]) <= c["max"], c
    r = A["vp8_random"]
    for e in r["amp"]:
        o3 = (C.c_int * 3)()
        L.orc_random_init(C.c_float(e["dithering"]), o3)
        assert (o3[0], o3[1], o3[2]) == (r["index1"], r["index2"], e["amp"]), e
    qq = A["quality_to_qindex"]
    assert L.orc_quality_to_qindex(0) == qq["q0"] and L.orc_quality_to_qindex(100) == qq["q100"]
    assert qq["q50_range"][0] <= L.orc_quality_to_qindex(50) <= qq["q50_range"][1]


def _check_quant(q, vin, out, nz):
    e = q["expect"]
    if "nz" in e:
        assert int(nz[0]) == e["nz"]
    if e.get("nz_positive"):
        assert int(nz[0]) > 0
    if "out0" in e:
        assert int(out[0, 0]) == e["out0"]
    if e.get("out0_equals_quantdiv"):  # (100 * iQ + B) >> 17 with iQ = (1 << 17) / 10, B = 96 << 9 (encode_test.go:168)
        assert abs(int(out[0, 0]) - ((int(vin[0, 0]) * ((1 << 17) // q["quant"]) + (q["bias_dc"] << 9)) >> 17)) <= 1
    if e.get("out1_nonzero"):
        assert int(out[0, 1]) != 0


def test_host_option_mapping_of_dithering():
    """The product's host mirror of the dithering amplitude (webp_b200/webp.py lossy_config: encode.go:563-567 derives the
    dithering strength 1 - 0.5 (Quality/100)^4 and dsp.InitRandom turns it into amp = int(256 * dithering), random.go:39-50):
    Quality 100 is the anchor's dithering 0.5 -> 128, Quality 0 its dithering 1.0 -> 256."""
    import webp_b200
    amp = {e["dithering"]: e["amp"] for e in A["vp8_random"]["amp"]}
    for quality, dithering in ((100, 0.5), (0, 1.0)):
        o = webp_b200.DefaultOptions()
        o.Quality = quality
        o.Preprocessing = 2
        assert webp_b200.webp.lossy_config(o).dither_amp == amp[dithering], quality


@pytest.mark.gpu
def test_gpu_twins_against_reference_anchors(gpu_ctx):
    from webp_b200 import dsp
    d, ty, by, tu, tv, bu, bv = _diamond_inputs()
    for ch in (3, 4):
        td, bd = dsp.UpsampleLinePairBatch(ty, by, tu, tv, bu, bv, channels=ch, ctx=gpu_ctx)
        _check_diamond(td, bd, d)
    for q in A["quantize"]:
        vin = np.array([q["in"]], np.int16)
        out, nz = dsp.QuantizeCoeffsBatch(vin, q["quant"], q["quant"], 0, 0, q["first"], gpu_ctx)
        _check_quant(q, vin, out, nz)
    dq = A["dequant"]
    out = dsp.DequantCoeffsBatch(np.array([dq["in"]], np.int16), dq["dc_quant"], dq["quant"], gpu_ctx)
    assert out[0, 0] == dq["expect"]["out0"] and out[0, 1] == dq["expect"]["out1"]
    for c in A["psnr"]["cases"]:
        assert c["min"] <= dsp.PSNRFromSSE(c["mse"], c["size"]) <= c["max"], c
