"""Animation frame path over the GPU lossy codec: the reference's animation package with its
per-frame codec calls batched.

Host-side mirror of (paths relative to the reference):
  AnimEncoder: NewEncoder / AddFrame / addOptimizedFrame / encodeKeyframe / encodeSubFrame / increasePreviousDuration /
    findChangedRect / snapToEven / isLossyBlendingPossible / sanitizeKeyframeOptions / Close       animation/animation.go:484-1218
  frame codec hooks: encodeFrameForAnimation / simpleEncodeForAnimation / decodeFrameForAnimation / ycbcrToNRGBA   webp.go:212-325
  Animation / DecodeBytes / DecodeFramesParallel / AnimDecoder (isKeyFrame, NextFrame, compositeFrame)  animation/animation.go:19-460
What the GPU changes: the reference encodes the candidates of a frame one after the other and decodes frames on a goroutine
pool; here the candidates of a frame go to the codec as ONE call, in all-keyframe mode (Kmax = 1) the whole clip is one
batch, and DecodeFrames decodes every frame of a file in one batch per frame size.
Outside this path (rejected, never emulated): Lossless / AllowMixed (VP8L), canvases with transparency and frames carrying
an ALPH chunk (the alpha plane is VP8L work).
"""
import math

import numpy as np

from . import mux
from . import webp

DisposeNone, DisposeBackground = 0, 1   # animation/frame.go:17-26
BlendAlpha, BlendNone = 0, 1            # animation/frame.go:28-36
maxDuration = mux.maxDuration
maxCanvasDimension = 16383


class AnimError(ValueError):
    pass


class EncodeOptions:
    """animation/animation.go:484."""

    def __init__(self, LoopCount=0, BackgroundColor=(0, 0, 0, 0), Quality=0, Lossless=False, AllowMixed=False, Kmin=0, Kmax=0):
        self.LoopCount, self.BackgroundColor, self.Quality, self.Lossless = LoopCount, tuple(BackgroundColor), Quality, Lossless
        self.AllowMixed, self.Kmin, self.Kmax = AllowMixed, Kmin, Kmax


def sanitizeKeyframeOptions(kmin, kmax):
    """animation/animation.go:546."""
    MaxInt = (1 << 63) - 1
    if kmax <= 0:
        return MaxInt - 1, MaxInt
    if kmax == 1:
        return 0, 0
    if kmin >= kmax:
        kmin = kmax - 1
    else:
        lim = kmax // 2 + 1
        if kmin < lim < kmax:
            kmin = lim
    if kmax - kmin > 30:
        kmin = kmax - 30
    return kmin, kmax


def nrgbaToARGB(c):
    r, g, b, a = c
    return (a << 24) | (r << 16) | (g << 8) | b


def argbToNRGBA(v):
    return ((v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF, (v >> 24) & 0xFF)


def _rect_intersect(r, s):
    x0, y0, x1, y1 = max(r[0], s[0]), max(r[1], s[1]), min(r[2], s[2]), min(r[3], s[3])
    return (x0, y0, x1, y1) if x0 < x1 and y0 < y1 else (0, 0, 0, 0)


def _rect_empty(r):
    return r[0] >= r[2] or r[1] >= r[3]


def findChangedRect(prev, curr):
    """animation/animation.go:1019: bounding rectangle of the pixels that differ ((0,0,0,0) when none does).  The reference's
    progressive narrowing finds the same extremes as the plain column test below."""
    diff = (prev != curr).any(axis=2)
    rows = np.flatnonzero(diff.any(axis=1))
    if rows.size == 0:
        return (0, 0, 0, 0)
    cols = np.flatnonzero(diff.any(axis=0))
    return (int(cols[0]), int(rows[0]), int(cols[-1]) + 1, int(rows[-1]) + 1)


def snapToEven(r):
    """animation/animation.go:1099."""
    w, h = r[2] - r[0] + (r[0] & 1), r[3] - r[1] + (r[1] & 1)
    x, y = r[0] & ~1, r[1] & ~1
    return (x, y, x + w, y + h)


def qualityToMaxDiff(quality):
    val = math.pow(quality / 100.0, 0.5)
    return int(31.0 * (1.0 - val) + 1.0 * val + 0.5)


def isLossyBlendingPossible(src, dst, rect, quality):
    """animation/animation.go:815: every pixel of rect is opaque in dst or close enough to src."""
    x0, y0, x1, y1 = rect
    s, d = src[y0:y1, x0:x1].astype(np.int32), dst[y0:y1, x0:x1].astype(np.int32)
    thr = qualityToMaxDiff(quality) * 255
    similar = (s[..., 3] == d[..., 3]) & ((np.abs(s[..., :3] - d[..., :3]) * d[..., 3:4]) <= thr).all(axis=2)
    return bool(((d[..., 3] == 0xFF) | similar).all())


def extractSubImage(src, rect):
    x0, y0, x1, y1 = rect
    if x1 - x0 <= 0 or y1 - y0 <= 0:
        return np.zeros((1, 1, 4), np.uint8)
    return np.ascontiguousarray(src[y0:y1, x0:x1])


def gpu_frame_encoder(quality, ctx=None):
    """encodeFrameForAnimation (webp.go:212) for a list of NRGBA frames: EncoderOptions{Quality, Method: 4}, every other field the
    Go zero value, raw VP8 bitstreams out.  Frames of one size go to the GPU as one batch."""
    opts = webp.EncoderOptions(Quality=float(quality), Method=4)

    def encode(frames):
        out = [None] * len(frames)
        groups = {}
        for i, f in enumerate(frames):
            groups.setdefault(f.shape[:2], []).append(i)
        for idx in groups.values():
            files = webp.EncodeBatch(np.stack([frames[i] for i in idx]), opts, ctx)
            for i, data in zip(idx, files):
                out[i] = mux.riff_payload(data)
        return out
    return encode


class AnimEncoder:
    """animation/animation.go:528.  frame_encoder(list of NRGBA uint8 arrays) -> list of raw VP8 bitstreams; the default is the
    GPU codec (the reference's FrameEncoderFunc hook, batched)."""

    def __init__(self, w, canvasWidth, canvasHeight, opts=None, frame_encoder=None, ctx=None):
        if canvasWidth <= 0 or canvasHeight <= 0 or canvasWidth > maxCanvasDimension or canvasHeight > maxCanvasDimension:
            raise AnimError("animation: invalid canvas dimensions")  # NewEncoder returns nil
        self.w, self.width, self.height = w, canvasWidth, canvasHeight
        o = opts or EncodeOptions()
        self.opts = EncodeOptions(min(max(o.LoopCount, 0), 0xFFFF), o.BackgroundColor, o.Quality, o.Lossless, o.AllowMixed, *sanitizeKeyframeOptions(o.Kmin, o.Kmax))
        if self.opts.Lossless or self.opts.AllowMixed:
            raise AnimError("animation: Lossless / AllowMixed (VP8L) are outside the GPU lossy path")
        self.muxer = mux.Muxer()
        self.muxer.SetCanvasSize(canvasWidth, canvasHeight)
        self.muxer.SetLoopCount(self.opts.LoopCount)
        self.muxer.SetBackgroundColor(nrgbaToARGB(self.opts.BackgroundColor))
        self.ctx = ctx
        self.encode = frame_encoder or gpu_frame_encoder(self.opts.Quality, ctx)
        self.closed = False
        self.prevCanvas = None
        self.frameCount = 0
        self.countSinceKeyframe = 0
        self.prevFrameRect = (0, 0, 0, 0)
        self.prevMuxIndex = 0
        self.codec_calls = 0  # how many times the frame codec was entered (one call may carry several frames)

    # -- helpers
    def _canvas(self, img):
        a = np.asarray(img)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] not in (3, 4):
            raise AnimError("animation: frames must be uint8 [h][w][3|4]")
        if a.shape[2] == 3:
            a = np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], axis=2)
        if a.shape[0] != self.height or a.shape[1] != self.width:  # smaller images sit at (0, 0) of a full canvas
            full = np.zeros((self.height, self.width, 4), np.uint8)
            hh, ww = min(a.shape[0], self.height), min(a.shape[1], self.width)
            full[:hh, :ww] = a[:hh, :ww]
            a = full
        if not bool((a[..., 3] == 255).all()):
            raise AnimError("animation: canvases with transparency need the ALPH (VP8L) chunk, which is outside the GPU lossy path")
        return np.ascontiguousarray(a)

    def _encode(self, frames):
        self.codec_calls += 1
        return self.encode(frames)

    def _commit_keyframe(self, canvas, bs, durMS):
        self.muxer.AddFrame(bs, mux.FrameOptions(Duration=durMS, BlendMode=BlendNone, DisposeMode=DisposeNone))
        self.prevCanvas = canvas.copy()
        self.prevFrameRect = (0, 0, self.width, self.height)
        self.prevMuxIndex = self.muxer.NumFrames() - 1
        self.frameCount += 1
        self.countSinceKeyframe = 0

    def increasePreviousDuration(self, durMS):
        """animation/animation.go:974."""
        new = self.muxer.FrameDuration(self.prevMuxIndex) + durMS
        if new < maxDuration:
            self.muxer.SetFrameDuration(self.prevMuxIndex, new)
            return
        # the reference emits a 1x1 transparent filler frame here: it needs an alpha plane
        raise AnimError("animation: merged duration overflows 24 bits; the transparent filler frame needs the ALPH (VP8L) chunk")

    # -- the reference's entry points
    def AddFrame(self, img, duration_ms):
        """animation/animation.go:616 + addOptimizedFrame :660 (durations in milliseconds)."""
        if self.closed:
            raise AnimError("animation: encoder is closed")
        canvas = self._canvas(img)
        durMS = int(duration_ms)
        if self.frameCount == 0:
            return self._commit_keyframe(canvas, self._encode([canvas])[0], durMS)
        if np.array_equal(self.prevCanvas, canvas):
            return self.increasePreviousDuration(durMS)
        self.countSinceKeyframe += 1
        if self.countSinceKeyframe >= self.opts.Kmax:
            return self._commit_keyframe(canvas, self._encode([canvas])[0], durMS)
        return self._encodeSubFrame(canvas, durMS)

    def _encodeSubFrame(self, curr, durMS):
        """animation/animation.go:846: dispose-none and dispose-background candidates of the frame, the smaller one wins; a
        full-canvas keyframe is tried when the winner covers more than 90 % of the canvas."""
        bounds = (0, 0, self.width, self.height)
        rectNone = findChangedRect(self.prevCanvas, curr)
        if _rect_empty(rectNone):
            rectNone = (0, 0, 1, 1)
        rectNone = _rect_intersect(snapToEven(rectNone), bounds)
        blendNone = BlendAlpha if isLossyBlendingPossible(self.prevCanvas, curr, rectNone, self.opts.Quality) else BlendNone
        disposed = self.prevCanvas.copy()
        px0, py0, px1, py1 = _rect_intersect(self.prevFrameRect, bounds)
        disposed[py0:py1, px0:px1] = 0
        rectBG = findChangedRect(disposed, curr)
        if _rect_empty(rectBG):
            rectBG = (0, 0, 1, 1)
        rectBG = _rect_intersect(snapToEven(rectBG), bounds)
        blendBG = BlendAlpha if isLossyBlendingPossible(disposed, curr, rectBG, self.opts.Quality) else BlendNone
        bsNone, bsBG = self._encode([extractSubImage(curr, rectNone), extractSubImage(curr, rectBG)])  # both candidates, one codec call
        useBG = len(bsBG) < len(bsNone)
        bestBS, bestRect, bestDispose, bestBlend = (bsBG, rectBG, DisposeBackground, blendBG) if useBG else (bsNone, rectNone, DisposeNone, blendNone)
        if (bestRect[2] - bestRect[0]) * (bestRect[3] - bestRect[1]) > self.width * self.height * 9 // 10:
            bsKey = self._encode([curr])[0]
            if len(bsKey) < len(bestBS):
                return self._commit_keyframe(curr, bsKey, durMS)  # encodeKeyframe encodes the same picture again: same bytes
        if bestDispose == DisposeBackground:
            self.muxer.SetFrameDisposeMode(self.prevMuxIndex, mux.DisposeBackground)
        self.muxer.AddFrame(bestBS, mux.FrameOptions(Duration=durMS, OffsetX=bestRect[0], OffsetY=bestRect[1], BlendMode=bestBlend, DisposeMode=DisposeNone))
        self.prevCanvas = curr.copy()
        self.prevFrameRect = bestRect
        self.prevMuxIndex = self.muxer.NumFrames() - 1
        self.frameCount += 1

    def AddFrames(self, imgs, durations_ms):
        """Batch twin of AddFrame.  With Kmax = 1 every frame is a full-canvas keyframe whose bytes depend on its own pixels only,
        so all frames that differ from their predecessor are encoded by ONE codec call; the muxer then sees exactly the
        AddFrame sequence.  Otherwise frames go through AddFrame one by one (their rectangles depend on the winners before)."""
        if len(imgs) != len(durations_ms):
            raise AnimError("animation: one duration per frame")
        if self.opts.Kmax != 0 or self.frameCount != 0:
            for img, d in zip(imgs, durations_ms):
                self.AddFrame(img, d)
            return
        if self.closed:
            raise AnimError("animation: encoder is closed")
        canvases = [self._canvas(img) for img in imgs]
        keep = [i for i in range(len(canvases)) if i == 0 or not np.array_equal(canvases[i - 1], canvases[i])]
        coded = dict(zip(keep, self._encode([canvases[i] for i in keep]))) if keep else {}
        for i, (canvas, d) in enumerate(zip(canvases, durations_ms)):
            if i in coded:
                self.countSinceKeyframe += 1 if self.frameCount else 0
                self._commit_keyframe(canvas, coded[i], int(d))
            else:
                self.increasePreviousDuration(int(d))

    def AddRawFrame(self, bitstream, duration_ms, offsetX=0, offsetY=0, blend=BlendAlpha, dispose=DisposeNone):
        if self.closed:
            raise AnimError("animation: encoder is closed")
        self.muxer.AddFrame(bitstream, mux.FrameOptions(Duration=int(duration_ms), OffsetX=offsetX, OffsetY=offsetY, BlendMode=blend, DisposeMode=dispose))

    def SetICCProfile(self, data):
        self.muxer.SetICCProfile(data)

    def SetEXIF(self, data):
        self.muxer.SetEXIF(data)

    def SetXMP(self, data):
        self.muxer.SetXMP(data)

    def Close(self, simple_encode=None):
        """animation/animation.go:1190: assemble; a one-frame animation is written as a still image when that file is smaller
        (simpleEncodeForAnimation, webp.go:229)."""
        if self.closed:
            return
        self.closed = True
        data = self.muxer.Assemble()
        if self.frameCount == 1 and self.prevCanvas is not None:
            if simple_encode is None:
                def simple_encode(canvas):
                    return webp.EncodeBatch(canvas[None], webp.EncoderOptions(Quality=float(self.opts.Quality), Method=4), self.ctx)[0]
            simple = simple_encode(self.prevCanvas)
            if 0 < len(simple) < len(data):
                data = simple
        self.w.write(data)


# ---------------------------------------------------------------------------------------------------------------- decode
class Frame:
    """animation/frame.go:38."""

    def __init__(self, fi):
        self.Image = None
        self.Duration, self.OffsetX, self.OffsetY = fi.Duration, fi.OffsetX, fi.OffsetY
        self.Dispose, self.Blend, self.IsKeyframe, self.HasAlpha = fi.DisposeMode, fi.BlendMode, fi.IsKeyframe, fi.HasAlpha
        self.BitstreamData, self.AlphaData = fi.Data, fi.AlphaData

    def Bounds(self):
        h, w = self.Image.shape[:2] if self.Image is not None else (0, 0)
        return (self.OffsetX, self.OffsetY, self.OffsetX + w, self.OffsetY + h)


class Animation:
    """animation/animation.go:19."""

    def __init__(self):
        self.CanvasWidth = self.CanvasHeight = 0
        self.LoopCount = 0
        self.BackgroundColor = (0, 0, 0, 0)
        self.Frames = []
        self.ICC = self.EXIF = self.XMP = None

    def TotalDuration(self):
        return sum(f.Duration for f in self.Frames)

    def DecodeFrames(self, ctx=None):
        """DecodeFrames / DecodeFramesParallel (animation/animation.go:164-258) with decodeFrameForAnimation (webp.go:243) as the
        frame codec: every undecoded frame of the file goes to the GPU decoder in one batch per frame size; planes come back and
        are converted by ycbcrToNRGBA."""
        todo = {}
        for i, f in enumerate(self.Frames):
            if f.Image is not None or not f.BitstreamData:
                continue
            if f.AlphaData or f.BitstreamData[0] == mux.VP8LMagicByte:
                raise AnimError("animation: frame %d carries an ALPH chunk or a VP8L bitstream: outside the GPU lossy path" % i)
            todo.setdefault(mux.parseVP8Dimensions(f.BitstreamData), []).append(i)
        for idx in todo.values():
            planes = webp.DecodeBatch([mux.writeRIFFSimple(mux.FourCCVP8, self.Frames[i].BitstreamData) for i in idx], ctx=ctx)
            for i, p in zip(idx, planes):
                self.Frames[i].Image = ycbcrToNRGBA(p.Y, p.Cb, p.Cr)

    DecodeFramesParallel = DecodeFrames


def ycbcrToNRGBA(y, cb, cr):
    """webp.go:272: nearest-neighbour 4:2:0 -> NRGBA with the JFIF constants (animation compositing does not use the fancy
    upsampler), int32 arithmetic with arithmetic shifts as in Go."""
    h, w = y.shape
    yy = y.astype(np.int32)
    cbf = np.repeat(np.repeat(cb.astype(np.int32) - 128, 2, axis=0), 2, axis=1)[:h, :w]
    crf = np.repeat(np.repeat(cr.astype(np.int32) - 128, 2, axis=0), 2, axis=1)[:h, :w]
    out = np.empty((h, w, 4), np.uint8)
    out[..., 0] = np.clip(yy + ((91881 * crf + 32768) >> 16), 0, 255)
    out[..., 1] = np.clip(yy - ((22554 * cbf + 46802 * crf + 32768) >> 16), 0, 255)
    out[..., 2] = np.clip(yy + ((116130 * cbf + 32768) >> 16), 0, 255)
    out[..., 3] = 255
    return out


def DecodeBytes(data):
    """animation/animation.go:104."""
    d = mux.Demuxer(data)
    a = Animation()
    a.CanvasWidth, a.CanvasHeight, a.LoopCount = d.Width, d.Height, d.LoopCount()
    a.BackgroundColor = argbToNRGBA(d.BackgroundColor())
    a.ICC, a.EXIF, a.XMP = d.iccData, d.exifData, d.xmpData
    a.Frames = [Frame(d.Frame(i)) for i in range(d.NumFrames())]
    return a


def Decode(r):
    return DecodeBytes(r if isinstance(r, (bytes, bytearray, memoryview)) else r.read())


def alphaBlendNRGBA(src, dst):
    """animation/animation.go:1243 on [..][4] uint8 arrays (non-premultiplied source over destination)."""
    s, d = src.astype(np.uint32), dst.astype(np.uint32)
    sa, da = s[..., 3], d[..., 3]
    dfa = (da * (256 - sa)) >> 8
    ba = sa + dfa
    scale = (1 << 24) // np.maximum(ba, 1)
    out = np.empty_like(src)
    for c in range(3):
        out[..., c] = np.minimum(((s[..., c] * sa + d[..., c] * dfa) * scale) >> 24, 255)
    out[..., 3] = ba
    out[ba == 0] = 0
    keep_dst = sa == 0
    take_src = ~keep_dst & ((sa == 255) | (da == 0))
    out[keep_dst] = dst[keep_dst]
    out[take_src] = src[take_src]
    return out


class AnimDecoder:
    """animation/animation.go:279: NextFrame composites the next frame and returns a snapshot of the canvas."""

    def __init__(self, anim):
        if anim.CanvasWidth <= 0 or anim.CanvasHeight <= 0:
            raise AnimError("animation: invalid canvas %dx%d" % (anim.CanvasWidth, anim.CanvasHeight))
        self.anim = anim
        self.currFrame = np.zeros((anim.CanvasHeight, anim.CanvasWidth, 4), np.uint8)
        self.prevFrameDisposed = np.zeros_like(self.currFrame)
        self.Reset()

    def Reset(self):
        self.pos = 0
        self.currFrame[:] = 0
        self.prevFrameDisposed[:] = 0
        self.prevFrameWasKeyframe = False
        self.prevDispose = DisposeNone
        self.prevBounds = (0, 0, 0, 0)

    def HasNext(self):
        return self.pos < len(self.anim.Frames)

    def isKeyFrame(self, idx):
        f = self.anim.Frames[idx]
        if idx == 0:
            return True
        cw, ch = self.anim.CanvasWidth, self.anim.CanvasHeight
        b = f.Bounds()
        if f.OffsetX == 0 and f.OffsetY == 0 and b[2] - b[0] == cw and b[3] - b[1] == ch and (not f.HasAlpha or f.Blend == BlendNone):
            return True
        if self.prevDispose == DisposeBackground:
            p = self.prevBounds
            if (p[0] == 0 and p[1] == 0 and p[2] - p[0] == cw and p[3] - p[1] == ch) or self.prevFrameWasKeyframe:
                return True
        return False

    def NextFrame(self):
        if not self.HasNext():
            raise AnimError("animation: no frames")
        f = self.anim.Frames[self.pos]
        if f.Image is None:
            raise AnimError("animation: frame image is nil")
        key = self.isKeyFrame(self.pos)
        if key:
            self.currFrame[:] = 0
        else:
            self.currFrame[:] = self.prevFrameDisposed
        bounds = (0, 0, self.anim.CanvasWidth, self.anim.CanvasHeight)
        x0, y0, x1, y1 = _rect_intersect(f.Bounds(), bounds)
        if x1 > x0:
            src = f.Image[y0 - f.OffsetY:y1 - f.OffsetY, x0 - f.OffsetX:x1 - f.OffsetX]
            dst = self.currFrame[y0:y1, x0:x1]
            dst[:] = src if f.Blend == BlendNone else alphaBlendNRGBA(src, dst)
        snap = self.currFrame.copy()
        self.prevFrameDisposed[:] = self.currFrame
        if f.Dispose == DisposeBackground:
            self.prevFrameDisposed[y0:y1, x0:x1] = 0
        self.prevFrameWasKeyframe, self.prevDispose, self.prevBounds = key, f.Dispose, f.Bounds()
        self.pos += 1
        return snap, f.Duration
