"""ctypes binding of libwebpgpu.so -- the same C ABI (include/webpgpu.h) the reference's cgo shim binds.

There is no CPU fallback: if the library is missing it is (re)built with nvcc, and every compute call
needs a CUDA device (wgpu_ctx_create fails loudly otherwise).
"""
import ctypes as C
import os
import re

from . import build as _build

_LIB = None

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM, ERR_BITSTREAM, ERR_TOO_SMALL = 0, -1, -2, -3, -4, -5, -6


class EncOptions(C.Structure):
    """wgpu_enc_options == lossy.EncodeConfig (internal/lossy/encode.go:46-63)."""
    _fields_ = [(n, C.c_int) for n in (
        "quality", "method", "sns_strength", "filter_strength", "filter_sharpness", "filter_type",
        "partitions", "segments", "preprocessing", "has_alpha", "passes", "dither_amp", "target_size")] + [
        ("target_psnr", C.c_float), ("qmin", C.c_int), ("qmax", C.c_int), ("use_sharp_yuv", C.c_int)]


class SegQuant(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("quant", "iquant", "bias", "dc_quant", "dc_iquant", "dc_bias")] + [("sharpen", C.c_int16 * 16)]


class Segment(C.Structure):
    """wgpu_segment == the SegmentInfo subset the mode search reads (internal/lossy/encode.go:278-323)."""
    _fields_ = [("y1", SegQuant), ("y2", SegQuant), ("uv", SegQuant)] + [(n, C.c_int) for n in (
        "lambda_i4", "lambda_i16", "lambda_uv", "lambda_mode", "tlambda_i4", "tlambda_i16", "tlambda_sd", "reserved")]


class WebPGPUError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (wgpu status %d)" % (msg, code))
        self.code = code


def header_symbols():
    """Every function name include/webpgpu.h declares."""
    path = os.path.join(os.path.dirname(_build.HERE), "include", "webpgpu.h")
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wgpu_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _LIB
    if _LIB is None:
        # WGPU_LIB: a differently built libwebpgpu.so (profiling builds, tools/phase_clock.py); same ABI, still no fallback
        path = os.environ.get("WGPU_LIB") or _build.build_native()
        L = C.CDLL(path)
        vp, u8p, i16p, i32p, sz = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t
        L.wgpu_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.wgpu_ctx_destroy.argtypes = [vp]
        L.wgpu_ctx_destroy.restype = None
        L.wgpu_last_error.argtypes = [vp]
        L.wgpu_last_error.restype = C.c_char_p
        L.wgpu_sync.argtypes = [vp]
        L.wgpu_set_host_threads.argtypes = [vp, C.c_int]
        L.wgpu_host_alloc.argtypes = [vp, sz]
        L.wgpu_host_alloc.restype = vp
        L.wgpu_host_free.argtypes = [vp, vp]
        L.wgpu_host_free.restype = None
        L.wgpu_pool_bucket.argtypes = [sz]
        L.wgpu_pool_bucket.restype = sz
        L.wgpu_ctx_mem_info.argtypes = [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(C.c_int)]
        L.wgpu_ctx_trim.argtypes = [vp]
        L.wgpu_enc_options_default.argtypes = [C.POINTER(EncOptions), C.c_int]
        L.wgpu_enc_options_default.restype = None
        L.wgpu_encode_batch.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, sz, C.POINTER(EncOptions), u8p, sz, vp]
        L.wgpu_enc_upload.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, sz]
        L.wgpu_enc_device.argtypes = [vp, C.POINTER(EncOptions)]
        L.wgpu_setup_segment.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Segment)]
        L.wgpu_enc_analyze.argtypes = [vp, C.POINTER(EncOptions), u8p, vp]
        L.wgpu_enc_search.argtypes = [vp, vp, u8p]
        L.wgpu_enc_finish.argtypes = [vp, u8p, sz, vp]
        L.wgpu_enc_fetch.argtypes = [vp, C.c_int] + [vp] * 11
        L.wgpu_decode_info.argtypes = [u8p, sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.wgpu_decode_batch.argtypes = [vp, vp, vp, C.c_int, u8p, u8p, u8p, sz, sz, u8p, sz]
        L.wgpu_dec_parse.argtypes = [vp, vp, vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.wgpu_dec_device.argtypes = [vp, C.c_int]
        L.wgpu_dec_fetch.argtypes = [vp, u8p, u8p, u8p, sz, sz, u8p, sz]
        L.wgpu_import_rgba.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, sz, C.c_int, u8p, u8p, u8p]
        L.wgpu_upsample_nrgba.argtypes = [vp, C.c_int, C.c_int, C.c_int, u8p, C.c_int, u8p, u8p, C.c_int, sz, sz, u8p, u8p]
        L.wgpu_plane_metrics.argtypes = [vp, C.c_int, u8p, u8p, C.c_int, C.c_int, C.c_int, sz, vp, vp]
        L.wgpu_psnr_from_sse.argtypes = [C.c_uint64, C.c_uint64]
        L.wgpu_psnr_from_sse.restype = C.c_double
        L.wgpu_dsp_ftransform_batch.argtypes = [vp, C.c_int, u8p, u8p, i16p]
        L.wgpu_dsp_itransform_batch.argtypes = [vp, C.c_int, u8p, i16p, u8p]
        L.wgpu_dsp_fwht_batch.argtypes = [vp, C.c_int, i16p, i16p]
        L.wgpu_dsp_iwht_batch.argtypes = [vp, C.c_int, i16p, i16p]
        L.wgpu_dsp_sse4x4_batch.argtypes = [vp, C.c_int, u8p, u8p, i32p]
        L.wgpu_dsp_tdisto4x4_batch.argtypes = [vp, C.c_int, u8p, u8p, i32p]
        L.wgpu_dsp_pred4_batch.argtypes = [vp, C.c_int, u8p, u8p]
        L.wgpu_dsp_pred_square_batch.argtypes = [vp, C.c_int, C.c_int, u8p, u8p]
        L.wgpu_dsp_quantize_batch.argtypes = [vp, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i16p, i32p]
        L.wgpu_dsp_trellis_batch.argtypes = [vp, C.c_int, i16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int, i16p, i32p]
        L.wgpu_dsp_token_cost_batch.argtypes = [vp, C.c_int, i16p, i32p, C.c_int, i32p, C.c_int, i32p]
        L.wgpu_dsp_sse16x16_batch.argtypes = [vp, C.c_int, u8p, u8p, i32p]
        L.wgpu_dsp_tdisto16x16_batch.argtypes = [vp, C.c_int, u8p, u8p, i32p]
        L.wgpu_dsp_dequant_batch.argtypes = [vp, C.c_int, i16p, C.c_int, C.c_int, i16p]
        L.wgpu_dsp_ftransform2_batch.argtypes = [vp, C.c_int, u8p, u8p, i16p]
        L.wgpu_dsp_dec_transform_batch.argtypes = [vp, C.c_int, C.c_int, i16p, u8p, u8p]
        L.wgpu_dsp_filter_batch.argtypes = [vp, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.wgpu_dsp_upsample_line_pair_batch.argtypes = [vp, C.c_int, C.c_int] + [u8p] * 8 + [C.c_int, u8p, u8p]
        L.wgpu_dsp_boolcode_batch.argtypes = [vp, C.c_int, vp, vp, u8p, sz, vp, C.POINTER(C.c_int)]
        L.wgpu_dec_reconstruct.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, u8p, C.c_int]
        L.wgpu_enc_stats.argtypes = [vp, C.c_int, vp]
        L.wgpu_cleanup_transparent.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, sz, u8p]
        L.wgpu_timer_begin.argtypes = [vp]
        L.wgpu_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
        L.wgpu_launch_count.argtypes = [vp]
        L.wgpu_launch_count.restype = C.c_uint64
        L.wgpu_transfer_bytes.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
        L.wgpu_enc_stage_time.argtypes = [vp, C.POINTER(EncOptions), C.c_int, C.c_int, C.POINTER(C.c_float)]
        _LIB = L
    return _LIB


class Context:
    """One GPU + one stream (wgpu_ctx).  Calls on one Context are serialised by the library."""

    def __init__(self, device=0, host_threads=0):
        self._h = C.c_void_p()
        rc = lib().wgpu_ctx_create(device, C.byref(self._h))
        if rc != OK:
            raise WebPGPUError(rc, (lib().wgpu_last_error(None) or b"wgpu_ctx_create failed").decode())
        self.device = device
        if host_threads:
            self.check(lib().wgpu_set_host_threads(self._h, host_threads))

    @property
    def handle(self):
        return self._h

    def check(self, rc):
        if rc != OK:
            raise WebPGPUError(rc, (lib().wgpu_last_error(self._h) or b"").decode())

    def launch_count(self):
        return int(lib().wgpu_launch_count(self._h))

    def transfer_bytes(self, reset=False):
        """(host->device, device->host) bytes copied by this context since the last reset."""
        a, b = C.c_uint64(), C.c_uint64()
        self.check(lib().wgpu_transfer_bytes(self._h, C.byref(a), C.byref(b), 1 if reset else 0))
        return int(a.value), int(b.value)

    def mem_info(self):
        """(device bytes, pinned host bytes, live buffers) this context's allocator holds (wgpu_ctx_mem_info)."""
        d, h, k = C.c_size_t(), C.c_size_t(), C.c_int()
        self.check(lib().wgpu_ctx_mem_info(self._h, C.byref(d), C.byref(h), C.byref(k)))
        return int(d.value), int(h.value), int(k.value)

    def trim(self):
        """Give the working buffers back (wgpu_ctx_trim); the next call re-reserves."""
        self.check(lib().wgpu_ctx_trim(self._h))

    def close(self):
        if self._h:
            lib().wgpu_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DEFAULT_CTX = {}


def default_context(device=0):
    if device not in _DEFAULT_CTX:
        _DEFAULT_CTX[device] = Context(device)
    return _DEFAULT_CTX[device]
