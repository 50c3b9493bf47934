"""CPU suite: the C-ABI library builds, loads and exports every symbol include/webpgpu.h declares; the host
mirror of the reference API validates options like validateConfig (encode.go:259); nothing computes on CPU."""
import io

import numpy as np
import pytest

import webp_b200
from webp_b200 import native


def test_library_exports_every_declared_symbol():
    L = native.lib()
    syms = native.header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "libwebpgpu.so does not export %s" % s


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(native.WebPGPUError) as e:
        native.Context(0)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises((native.WebPGPUError, webp_b200.WebPError)):
        webp_b200.Encode(io.BytesIO(), np.zeros((64, 64, 4), np.uint8) + 255)


def test_product_never_imports_oracle():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "webp_b200")):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in text.replace("oracle/ ", "") or f == "build.py", f
                assert "liboracle" not in text and "oracle_lib" not in text, f


def test_default_options_and_mapping():
    o = webp_b200.DefaultOptions()
    assert (o.Quality, o.Method, o.SNSStrength, o.FilterStrength, o.Segments) == (75, 4, -1, -1, -1)
    c = webp_b200.webp.lossy_config(o)
    assert (c.quality, c.method, c.sns_strength, c.filter_strength, c.filter_type, c.segments, c.partitions) == (75, 4, 50, 60, 1, 4, 0)
    p = webp_b200.OptionsForPreset(webp_b200.webp.PresetText, 60)
    assert (p.SNSStrength, p.FilterStrength, p.Segments) == (0, 0, 2)
    d = native.EncOptions()
    native.lib().wgpu_enc_options_default(d, 75)
    assert [getattr(d, f) for f, _ in d._fields_] == [getattr(c, f) for f, _ in c._fields_]


def test_enc_options_layout_matches_the_header():
    """wgpu_enc_options in include/webpgpu.h, its ctypes mirror and the oracle's config struct list the same fields in the same order
    (a field added on one side only would silently shift every later option)."""
    import ctypes as C, os, re, sys
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import oracle_lib
    text = open(os.path.join(os.path.dirname(here), "include", "webpgpu.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} wgpu_enc_options;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            ctype, names = decl.split(None, 1)
            fields += [(n.strip(), ctype) for n in names.split(",")]
    mirror = [(n, "float" if t is C.c_float else "int") for n, t in native.EncOptions._fields_]
    assert fields == mirror
    assert [n for n, _ in oracle_lib.OrcEncCfg._fields_] == [n for n, _ in fields]
    assert C.sizeof(native.EncOptions) == 4 * len(fields)
    o = webp_b200.DefaultOptions()
    o.UseSharpYUV = True  # EncoderOptions.UseSharpYUV (encode.go:62) -> use_sharp_yuv
    assert webp_b200.webp.lossy_config(o).use_sharp_yuv == 1 and webp_b200.validateConfig(o) is None


def test_rate_control_option_mapping():
    # EncoderOptions -> lossy.EncodeConfig for the doSearch fields (encode.go:480-488): QMax -1 is the "unset" sentinel (resolveQMax)
    o = webp_b200.DefaultOptions()
    o.TargetSize, o.TargetPSNR, o.QMin, o.QMax = 12000, 41.5, 10, -1
    c = webp_b200.webp.lossy_config(o)
    assert (c.target_size, c.qmin, c.qmax) == (12000, 10, 100) and abs(c.target_psnr - 41.5) < 1e-6
    o.QMax = 80
    assert webp_b200.webp.lossy_config(o).qmax == 80
    o.QMin = 90
    assert "QMin/QMax" in webp_b200.validateConfig(o)


@pytest.mark.parametrize("field,value,frag", [
    ("Quality", 101, "invalid Quality"), ("Method", 7, "invalid Method"), ("FilterSharpness", 8, "invalid FilterSharpness"),
    ("Partitions", 4, "invalid Partitions"), ("Segments", 5, "invalid Segments"), ("SNSStrength", 101, "invalid SNSStrength"),
    ("Preprocessing", 4, "invalid Preprocessing"), ("TargetSize", -1, "invalid TargetSize"),
])
def test_validate_config_rejects(field, value, frag):
    o = webp_b200.DefaultOptions()
    setattr(o, field, value)
    assert frag in webp_b200.validateConfig(o)
    with pytest.raises(webp_b200.WebPError):
        webp_b200.Encode(io.BytesIO(), np.full((64, 64, 4), 255, np.uint8), o)


def test_nil_arguments():
    with pytest.raises(webp_b200.WebPError, match="nil writer"):
        webp_b200.Encode(None, np.zeros((64, 64, 4), np.uint8))
    with pytest.raises(webp_b200.WebPError, match="nil image"):
        webp_b200.Encode(io.BytesIO(), None)
    with pytest.raises(webp_b200.WebPError, match="nil reader"):
        webp_b200.Decode(None)


def test_decode_config_is_host_only(oracle):
    data = oracle.encode(oracle.synth_image(80, 64, 1))
    cfg = webp_b200.DecodeConfig(io.BytesIO(data))
    assert (cfg.Width, cfg.Height) == (80, 64)
    with pytest.raises(webp_b200.WebPError):
        webp_b200.DecodeConfig(b"RIFF\x00\x00\x00\x00WEBPnope")
