"""CPU model of the arithmetic used by boolcode_kernel (webp_b200/csrc/token_kernels.cuh), checked against a literal
restatement of VP8BitWriter PutBit / Flush / Finish (internal/bitio/writer_bool.go:58-150).  It pins the three identities the
kernel rests on -- not the CUDA code itself, which the GPU parity tests cover:
  1. the range recurrence in R = range + 1 form with one multiply-add per token (sx = (R * +-prob + c) >> 8);
  2. folding four tokens' (increment, shift) events into the pending value at once;
  3. emitting all complete bytes of a step together, with the carry applied to the byte held back, 0xff runs excepted."""
import random

import pytest


def reference_writer(tokens):
    rng, value, run, nb, out = 254, 0, 0, -8, bytearray()

    def flush():
        nonlocal value, nb, run
        s = 8 + nb
        bits = value >> s
        value -= bits << s
        nb -= 8
        if (bits & 0xff) != 0xff:
            if (bits & 0x100) and out:
                out[-1] = (out[-1] + 1) & 0xff
            if run > 0:
                out.extend([0x00 if bits & 0x100 else 0xff] * run)
                run = 0
            out.append(bits & 0xff)
        else:
            run += 1

    def put(bit, prob):
        nonlocal rng, value, nb
        split = (rng * prob) >> 8
        if bit:
            value += split + 1
            rng -= split + 1
        else:
            rng = split
        shift = (7 - ((rng + 1).bit_length() - 1)) if rng < 127 else 0
        rng = ((rng + 1) << shift) - 1
        value <<= shift
        nb += shift
        if nb > 0:
            flush()

    for b, p in tokens:
        put(b, p)
    for _ in range(9 - nb):
        put(0, 128)
    nb = 0
    flush()
    return bytes(out)


def kernel_model(tokens):
    # range warp: events (increment, shift) per token, then the 17 closing events
    R, evs = 255, []

    def step(bit, prob):
        nonlocal R
        ma, mc = (-prob, prob - 1) if bit else (prob, -prob)
        sx = (R * ma + mc) >> 8
        r1 = (R if bit else 1) + sx
        k = r1.bit_length() - 1
        R = (r1 << 7) >> k
        return (-sx if bit else 0), 7 - k

    evs = [step(b, p) for b, p in tokens]
    fin = [step(0, 128) for _ in range(17)]
    # byte warp
    value, run, nb, last, out = 0, 0, -8, -1, bytearray()

    def flush():
        nonlocal value, nb, run, last
        s = 8 + nb
        bits = value >> s
        value -= bits << s
        nb -= 8
        assert bits < 512
        if (bits & 0xff) != 0xff:
            c = (bits >> 8) & 1
            if last >= 0:
                out.append((last + c) & 0xff)
            if run > 0:
                out.extend([0 if c else 0xff] * run)
                run = 0
            last = bits & 0xff
        else:
            run += 1

    evs4 = evs + [(0, 0)] * (-len(evs) % 4)  # the kernel pads a chunk with no-op events
    for i in range(0, len(evs4), 4):
        (a0, s0), (a1, s1), (a2, s2), (a3, s3) = evs4[i:i + 4]
        S3 = s3; S2 = s2 + S3; S1 = s1 + S2; S0 = s0 + S1
        value = (value << S0) + (a0 << S0) + (a1 << S1) + (a2 << S2) + (a3 << S3)
        nb += S0
        assert value < (1 << 64)
        nbytes = (nb + 7) >> 3
        if nbytes > 0:
            s_low = 16 + nb - 8 * nbytes
            cb = value >> s_low
            b = cb & 0xffffffff
            ff = b & (b >> 1); f2 = ff & (ff >> 2); f4 = f2 & (f2 >> 4)
            mask = 0x01010101 if nbytes == 4 else ((1 << (8 * nbytes)) - 1) & 0x01010101
            if run > 0 or (f4 & mask):
                while nb > 0:
                    flush()
            else:
                c = (cb >> (8 * nbytes)) & 1
                assert cb >> (8 * nbytes + 1) == 0
                if last >= 0:
                    out.append((last + c) & 0xff)
                for j in range(nbytes - 1):
                    out.append((b >> (8 * (nbytes - 1 - j))) & 0xff)
                last = b & 0xff
                value -= cb << s_low
                nb -= 8 * nbytes
    for i in range(9 - nb):
        a, s = fin[i]
        value = (value + a) << s
        nb += s
        if nb > 0:
            flush()
    nb = 0
    flush()
    if last >= 0:
        out.append(last)
    return bytes(out)


@pytest.mark.parametrize("mode", range(5))
def test_kernel_arithmetic_equals_the_reference_writer(mode):
    rnd = random.Random(100 + mode)
    for trial in range(120):
        toks = []
        for _ in range(rnd.randint(0, 2500)):
            if mode == 0:
                p, b = rnd.randint(0, 255), rnd.randint(0, 1)
            elif mode == 1:
                p, b = rnd.choice([1, 2, 254, 255, 128]), rnd.randint(0, 1)
            elif mode == 2:
                p, b = rnd.randint(200, 255), int(rnd.random() < 0.9)   # long carry chains
            elif mode == 3:
                p, b = rnd.randint(0, 30), int(rnd.random() >= 0.1)
            else:
                p, b = rnd.choice([255, 254, 253]), 1                    # 0xff runs
            toks.append((b, p))
        assert kernel_model(toks) == reference_writer(toks), (mode, trial, len(toks))


def test_sub_range_is_one_multiply_add():
    # r + 1 = (R * a + c) >> 8 with (a, c) = (256 - p, p - 1) for a one, (p, 256 - p) for a zero
    for R in range(128, 256):
        for p in range(256):
            split = ((R - 1) * p) >> 8
            assert (R * (256 - p) + p - 1) >> 8 == R - 1 - split
            assert (R * p + 256 - p) >> 8 == split + 1
