"""RIFF / VP8X container around the lossy frames the GPU path produces and consumes.

Host-side mirror of the reference's container code for VP8 frames -- byte layout, chunk order, flag bits and error strings:
  Muxer (SetCanvasSize / SetLoopCount / SetBackgroundColor / AddFrame / Assemble)     mux/mux.go:14-681
  Demuxer (VP8X flags, ICCP / EXIF / XMP, ANIM, ANMF sub-chunks)                        mux/demux.go:13-589
  writeRIFFSimple / writeRIFFExtended (webp.Encode with EncoderOptions.ICC/EXIF/XMP)    encode.go:955-1121
In the Go drop-in this code is the reference's own (BASELINE.json north_star: "the container/mux code serialise them on
the host"); it is restated here because the host side above the C ABI is Python in this image (no Go toolchain).
What stays outside: ALPH payloads are carried through byte for byte but never produced or decoded (the alpha plane is VP8L
work), VP8L frames are muxed / demuxed as opaque byte strings.
"""
import struct

FourCCRIFF, FourCCWEBP = b"RIFF", b"WEBP"
FourCCVP8, FourCCVP8L, FourCCVP8X = b"VP8 ", b"VP8L", b"VP8X"
FourCCALPH, FourCCANIM, FourCCANMF = b"ALPH", b"ANIM", b"ANMF"
FourCCICCP, FourCCEXIF, FourCCXMP = b"ICCP", b"EXIF", b"XMP "
ChunkHeaderSize, RIFFHeaderSize, VP8XChunkSize, ANIMChunkSize, ANMFChunkSize = 8, 12, 10, 6, 16
MaxCanvasSize, MaxFrames = 1 << 24, 10000
maxDuration, maxLoopCount = 0xFFFFFF, 0xFFFF
VP8LMagicByte = 0x2F
flagAnimation, flagXMP, flagEXIF, flagAlpha, flagICCP = 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 5  # mux/demux.go:53-57
BlendAlpha, BlendNone = 0, 1            # mux/demux.go:16-17
DisposeNone, DisposeBackground = 0, 1   # mux/demux.go:24-25


class MuxError(ValueError):
    pass


def _le24(v):
    return bytes((v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF))


def _chunk(fourcc, data):
    """writeDataChunk (mux/mux.go:659): header + payload + one zero byte after an odd payload."""
    return fourcc + struct.pack("<I", len(data)) + bytes(data) + (b"\0" if len(data) & 1 else b"")


def chunkTotalSize(n):
    return ChunkHeaderSize + n + (n & 1)


def parseVP8Dimensions(data):
    """mux/demux.go:544."""
    if len(data) < 10 or data[3:6] != b"\x9d\x01\x2a":
        return 0, 0
    w, h = struct.unpack_from("<HH", data, 6)
    return w & 0x3FFF, h & 0x3FFF


def splitAlphaAndBitstream(data):
    """mux/mux.go:477: frame data may carry its ALPH chunk in front of the VP8 bitstream."""
    if len(data) >= ChunkHeaderSize and data[0:4] == FourCCALPH:
        n = struct.unpack_from("<I", data, 4)[0]
        end = ChunkHeaderSize + n
        if end <= len(data):
            rest = end + (1 if (n & 1) and end < len(data) else 0)
            return data[ChunkHeaderSize:end], data[rest:]
    return None, data


def frameDimensions(data):
    """mux/mux.go:605."""
    _, bs = splitAlphaAndBitstream(data)
    if len(bs) >= 5 and bs[0] == VP8LMagicByte:
        bits = struct.unpack_from("<I", bs, 1)[0]
        return (bits & 0x3FFF) + 1, ((bits >> 14) & 0x3FFF) + 1
    return parseVP8Dimensions(bs)


def detectBitstreamType(data):
    return FourCCVP8L if len(data) > 0 and data[0] == VP8LMagicByte else FourCCVP8


def riff_payload(file_bytes):
    """The raw VP8 bitstream inside a simple RIFF/WEBP file (what encodeLossy hands to writeRIFF, encode.go:968)."""
    if len(file_bytes) < 20 or file_bytes[0:4] != FourCCRIFF or file_bytes[8:12] != FourCCWEBP or file_bytes[12:16] != FourCCVP8:
        raise MuxError("mux: not a simple lossy WebP file")
    n = struct.unpack_from("<I", file_bytes, 16)[0]
    return bytes(file_bytes[20:20 + n])


def writeRIFFSimple(fourcc, bitstream):
    """encode.go:968-997."""
    n = len(bitstream)
    padded = n + (n & 1)
    return FourCCRIFF + struct.pack("<I", 4 + ChunkHeaderSize + padded) + FourCCWEBP + _chunk(fourcc, bitstream)


def writeRIFFExtended(fourcc, bitstream, alpha, width, height, icc=b"", exif=b"", xmp=b""):
    """encode.go:1002-1115: RIFF -> VP8X -> [ICCP] -> [ALPH] -> VP8/VP8L -> [EXIF] -> [XMP]."""
    flags = 0
    if alpha:
        flags |= 0x10
    if fourcc == FourCCVP8L and len(bitstream) >= 5 and bitstream[0] == VP8LMagicByte and (struct.unpack_from("<I", bitstream, 1)[0] >> 28) & 1:
        flags |= 0x10
    if icc:
        flags |= 0x20
    if exif:
        flags |= 0x08
    if xmp:
        flags |= 0x04
    body = FourCCVP8X + struct.pack("<II", VP8XChunkSize, flags) + _le24(width - 1) + _le24(height - 1)
    if icc:
        body += _chunk(FourCCICCP, icc)
    if alpha:
        body += _chunk(FourCCALPH, alpha)
    body += _chunk(fourcc, bitstream)
    if exif:
        body += _chunk(FourCCEXIF, exif)
    if xmp:
        body += _chunk(FourCCXMP, xmp)
    riff = 4 + len(body)
    if riff > 0xFFFFFFFF - 8:
        raise MuxError("webp: RIFF payload too large (%d bytes)" % riff)
    return FourCCRIFF + struct.pack("<I", riff) + FourCCWEBP + body


def writeRIFF(fourcc, bitstream, alpha, width, height, icc=b"", exif=b"", xmp=b""):
    """encode.go:955: the extended format only when an alpha plane or metadata asks for it."""
    if alpha or icc or exif or xmp:
        return writeRIFFExtended(fourcc, bitstream, alpha, width, height, icc, exif, xmp)
    return writeRIFFSimple(fourcc, bitstream)


class FrameOptions:
    """mux/mux.go:14."""

    def __init__(self, Duration=0, OffsetX=0, OffsetY=0, BlendMode=BlendAlpha, DisposeMode=DisposeNone):
        self.Duration, self.OffsetX, self.OffsetY, self.BlendMode, self.DisposeMode = Duration, OffsetX, OffsetY, BlendMode, DisposeMode


class Muxer:
    """mux/mux.go:28-681."""

    def __init__(self):
        self.frames = []  # [data, FrameOptions]
        self.iccData = self.exifData = self.xmpData = None
        self.bgColor = 0
        self.loopCount = 0
        self.canvasWidth = self.canvasHeight = 0

    def SetICCProfile(self, data):
        self.iccData = bytes(data)

    def SetEXIF(self, data):
        self.exifData = bytes(data)

    def SetXMP(self, data):
        self.xmpData = bytes(data)

    def SetBackgroundColor(self, color):
        self.bgColor = color & 0xFFFFFFFF

    def SetLoopCount(self, count):
        self.loopCount = min(max(count, 0), maxLoopCount)

    def SetCanvasSize(self, width, height):
        self.canvasWidth, self.canvasHeight = min(width, MaxCanvasSize), min(height, MaxCanvasSize)

    def AddFrame(self, data, opts=None):
        if len(data) == 0:
            raise MuxError("mux: frame data is empty")
        if len(self.frames) >= MaxFrames:
            raise MuxError("mux: too many frames (max %d)" % MaxFrames)
        o = opts or FrameOptions()
        fo = FrameOptions(min(max(o.Duration, 0), maxDuration), o.OffsetX, o.OffsetY, o.BlendMode, o.DisposeMode)
        self.frames.append([bytes(data), fo])

    def SetFrameDisposeMode(self, index, mode):
        if 0 <= index < len(self.frames):
            self.frames[index][1].DisposeMode = mode

    def SetFrameDuration(self, index, ms):
        if 0 <= index < len(self.frames):
            self.frames[index][1].Duration = min(max(ms, 0), maxDuration)

    def FrameDuration(self, index):
        return self.frames[index][1].Duration if 0 <= index < len(self.frames) else 0

    def NumFrames(self):
        return len(self.frames)

    def isAnimated(self):
        return len(self.frames) > 1 or any(f[1].Duration > 0 for f in self.frames)

    def needsVP8X(self):
        return self.isAnimated() or self.iccData is not None or self.exifData is not None or self.xmpData is not None

    def hasAlpha(self):
        for data, _ in self.frames:
            if len(data) >= 12 and data[0:4] == FourCCALPH:
                return True
            if len(data) >= 5 and data[0] == VP8LMagicByte and (struct.unpack_from("<I", data, 1)[0] >> 28) & 1:
                return True
        return False

    def canvasSize(self):
        if self.canvasWidth > 0 and self.canvasHeight > 0:
            return self.canvasWidth, self.canvasHeight
        mw = mh = 0
        for data, o in self.frames:
            fw, fh = frameDimensions(data)
            mw, mh = max(mw, o.OffsetX + fw), max(mh, o.OffsetY + fh)
        return mw or 1, mh or 1

    def validate(self):
        if not self.frames:
            raise MuxError("mux: no frames to assemble")
        if not self.isAnimated() and len(self.frames) != 1:
            raise MuxError("mux: validation failed: non-animated image must have exactly 1 frame")
        cw, ch = self.canvasSize()
        for i, (data, o) in enumerate(self.frames):
            fw, fh = frameDimensions(data)
            if fw == 0 or fh == 0:
                continue
            if o.OffsetX + fw > cw or o.OffsetY + fh > ch:
                raise MuxError("mux: validation failed: frame %d (%dx%d at %d,%d) exceeds canvas (%dx%d)" % (i, fw, fh, o.OffsetX, o.OffsetY, cw, ch))

    def Assemble(self):
        """mux/mux.go:219: returns the file's bytes (the reference writes them to an io.Writer)."""
        self.validate()
        if not self.needsVP8X():
            data = self.frames[0][0]
            return writeRIFFSimple(detectBitstreamType(data), data)
        animated = self.isAnimated()
        flags = (flagAnimation if animated else 0) | (flagICCP if self.iccData is not None else 0) | (flagEXIF if self.exifData is not None else 0)
        flags |= (flagXMP if self.xmpData is not None else 0) | (flagAlpha if self.hasAlpha() else 0)
        cw, ch = self.canvasSize()
        body = FourCCVP8X + struct.pack("<I", VP8XChunkSize) + bytes((flags, 0, 0, 0)) + _le24(cw - 1) + _le24(ch - 1)
        if self.iccData is not None:
            body += _chunk(FourCCICCP, self.iccData)
        if animated:
            body += FourCCANIM + struct.pack("<IIH", ANIMChunkSize, self.bgColor, self.loopCount & 0xFFFF)
        parts = [body]
        for data, o in self.frames:
            parts.append(self._anmf(data, o) if animated else _chunk(detectBitstreamType(data), data))
        if self.exifData is not None:
            parts.append(_chunk(FourCCEXIF, self.exifData))
        if self.xmpData is not None:
            parts.append(_chunk(FourCCXMP, self.xmpData))
        body = b"".join(parts)
        if 4 + len(body) > 0xFFFFFFFF:
            raise MuxError("mux: RIFF payload too large (%d bytes, exceeds 4GB limit)" % (4 + len(body)))
        return FourCCRIFF + struct.pack("<I", 4 + len(body)) + FourCCWEBP + body

    @staticmethod
    def _anmf(data, o):
        """writeANMFChunk (mux/mux.go:503): 16-byte frame header, [ALPH], VP8 / VP8L."""
        alpha, bs = splitAlphaAndBitstream(data)
        sub = (_chunk(FourCCALPH, alpha) if alpha is not None else b"") + _chunk(detectBitstreamType(bs), bs)
        fw, fh = frameDimensions(data)
        hdr = _le24(o.OffsetX // 2) + _le24(o.OffsetY // 2) + (_le24(fw - 1) + _le24(fh - 1) if fw > 0 and fh > 0 else bytes(6))
        hdr += _le24(o.Duration) + bytes(((1 if o.DisposeMode == DisposeBackground else 0) | (2 if o.BlendMode == BlendNone else 0),))
        payload = ANMFChunkSize + len(sub)
        return FourCCANMF + struct.pack("<I", payload) + hdr + sub + (b"\0" if payload & 1 else b"")


class FrameInfo:
    """mux/demux.go:73."""

    def __init__(self, **kw):
        self.Data = b""; self.AlphaData = b""; self.Width = self.Height = 0; self.OffsetX = self.OffsetY = 0; self.Duration = 0
        self.IsKeyframe = False; self.HasAlpha = False; self.BlendMode = BlendAlpha; self.DisposeMode = DisposeNone
        self.__dict__.update(kw)


def _read_chunk(buf, pos):
    """ReadChunk (mux/chunk.go:63): (fourcc, payload, bytes consumed) or None when the header or payload does not fit."""
    if pos + ChunkHeaderSize > len(buf):
        return None
    n = struct.unpack_from("<I", buf, pos + 4)[0]
    if pos + ChunkHeaderSize + n > len(buf):
        return None
    adv = ChunkHeaderSize + n
    if (n & 1) and pos + adv < len(buf):
        adv += 1
    return bytes(buf[pos:pos + 4]), bytes(buf[pos + 8:pos + 8 + n]), adv


def _frame_has_alpha(data):
    return len(data) >= 5 and data[0] == VP8LMagicByte and bool((struct.unpack_from("<I", data, 1)[0] >> 28) & 1)


class Demuxer:
    """mux/demux.go:88-540 (NewDemuxer parses on construction)."""

    def __init__(self, data):
        self.data = bytes(data)
        self.frames = []
        self.iccData = self.exifData = self.xmpData = None
        self.bgColor = 0
        self.loopCount = 0
        self.Width = self.Height = 0
        self.HasAlpha = self.HasAnimation = self.HasICC = self.HasEXIF = self.HasXMP = False
        self.Format = 0  # 1 lossy, 2 lossless, 3 extended (mux/demux.go:31-36)
        self._parse()

    def NumFrames(self):
        return len(self.frames)

    def Frame(self, i):
        if not 0 <= i < len(self.frames):
            raise MuxError("mux: frame index out of range")
        return self.frames[i]

    def LoopCount(self):
        return self.loopCount

    def BackgroundColor(self):
        return self.bgColor

    def _parse(self):
        d = self.data
        if len(d) < RIFFHeaderSize or d[0:4] != FourCCRIFF or d[8:12] != FourCCWEBP:
            raise MuxError("mux: not a valid WebP file (bad RIFF header)")
        total = min(struct.unpack_from("<I", d, 4)[0] + 8, len(d))  # truncated data: work with what is there
        payload = d[RIFFHeaderSize:total]
        if len(payload) < ChunkHeaderSize:
            raise MuxError("mux: no image data found")
        first = payload[0:4]
        if first == FourCCVP8X:
            return self._parse_extended(payload)
        if first in (FourCCVP8, FourCCVP8L):
            c = _read_chunk(payload, 0)
            if c is None:
                raise MuxError("mux: data truncated")
            w, h = frameDimensions(c[1])
            if w == 0:
                raise MuxError("mux: invalid frame bitstream")
            self.Width, self.Height, self.Format = w, h, 1 if first == FourCCVP8 else 2
            self.HasAlpha = _frame_has_alpha(c[1])
            self.frames.append(FrameInfo(Data=c[1], Width=w, Height=h, IsKeyframe=True, HasAlpha=self.HasAlpha))
            return
        raise MuxError("mux: unknown first chunk %s" % first.decode("latin1"))

    def _parse_extended(self, payload):
        c = _read_chunk(payload, 0)
        if c is None:
            raise MuxError("mux: data truncated")
        if len(c[1]) < VP8XChunkSize:
            raise MuxError("mux: invalid VP8X chunk")
        x = c[1]
        flags = x[0]
        self.Width = (x[4] | x[5] << 8 | x[6] << 16) + 1
        self.Height = (x[7] | x[8] << 8 | x[9] << 16) + 1
        self.HasAlpha, self.HasAnimation = bool(flags & flagAlpha), bool(flags & flagAnimation)
        self.HasICC, self.HasEXIF, self.HasXMP = bool(flags & flagICCP), bool(flags & flagEXIF), bool(flags & flagXMP)
        self.Format = 3
        pos = c[2]
        pending_alpha = None
        while pos + ChunkHeaderSize <= len(payload):
            c = _read_chunk(payload, pos)
            if c is None:
                break
            fcc, data, adv = c
            if fcc == FourCCICCP:
                self.iccData = data
            elif fcc == FourCCEXIF:
                self.exifData = data
            elif fcc == FourCCXMP:
                self.xmpData = data
            elif fcc == FourCCANIM:
                if len(data) < ANIMChunkSize:
                    raise MuxError("mux: invalid ANIM chunk")
                self.bgColor, self.loopCount = struct.unpack_from("<IH", data, 0)
            elif fcc == FourCCANMF:
                self._parse_anmf(data)
            elif fcc == FourCCALPH and not self.HasAnimation and not self.frames:
                pending_alpha = data
            elif fcc in (FourCCVP8, FourCCVP8L) and not self.HasAnimation and not self.frames:
                w, h = frameDimensions(data)
                self.frames.append(FrameInfo(Data=data, AlphaData=pending_alpha or b"", Width=w, Height=h, IsKeyframe=True,
                                             HasAlpha=bool(pending_alpha) or _frame_has_alpha(data)))
            pos += adv
        if not self.frames:
            raise MuxError("mux: no image data found")

    def _parse_anmf(self, data):
        """mux/demux.go:401."""
        if len(data) < ANMFChunkSize:
            raise MuxError("mux: invalid ANMF chunk")
        le24 = lambda o: data[o] | data[o + 1] << 8 | data[o + 2] << 16
        image = alpha = b""
        sub = data[ANMFChunkSize:]
        pos = 0
        while pos + ChunkHeaderSize <= len(sub):
            c = _read_chunk(sub, pos)
            if c is None:
                break
            if c[0] in (FourCCVP8, FourCCVP8L):
                image = c[1]
            elif c[0] == FourCCALPH:
                alpha = c[1]
            pos += c[2]
        if len(self.frames) >= MaxFrames:
            raise MuxError("mux: too many frames: exceeded limit of %d" % MaxFrames)
        self.frames.append(FrameInfo(Data=image, AlphaData=alpha, OffsetX=le24(0) * 2, OffsetY=le24(3) * 2, Width=le24(6) + 1, Height=le24(9) + 1,
                                     Duration=le24(12), IsKeyframe=not self.frames, HasAlpha=len(alpha) > 0 or _frame_has_alpha(image),
                                     BlendMode=BlendNone if data[15] & 2 else BlendAlpha, DisposeMode=DisposeBackground if data[15] & 1 else DisposeNone))
