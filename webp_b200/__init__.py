"""webp_b200: B200-native (sm_100a) VP8 lossy pixel pipeline behind deepteams/webp's Encode/Decode/dsp surface.

The product is webp_b200/csrc (CUDA kernels + the C ABI of include/webpgpu.h, built into
webp_b200/_build/libwebpgpu.so); this package is the host-side mirror of the reference's Go interface.
"""
from .webp import (Config, Decode, DecodeBatch, DecodeConfig, DefaultOptions, Encode, EncodeBatch, EncoderOptions,  # noqa: F401
                   Options, OptionsForPreset, WebPError, YCbCr, validateConfig)
from . import animation, dsp, mux, native  # noqa: F401
