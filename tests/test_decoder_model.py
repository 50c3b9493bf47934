"""CPU model of the boolean-decoder hand-over used by the GPU macroblock parser: the host parses the frame headers with a
56-bit-refill decoder (host_dec.h::BoolDec), rewinds its state to a byte boundary and the device continues with a 32-bit
window refilled 16 bits at a time (dec_parse.cuh::DBoolDec).  Both must read the same bits and hit end-of-data at the same
point whatever the hand-over position (bitio/reader_bool.go semantics: one virtual zero byte past the end, then EOF)."""
import random


class HostDec:  # wgh::BoolDec
    def __init__(self, data):
        self.d, self.p, self.value, self.range, self.bits, self.eof = data, 0, 0, 254, -8, False
        self.refill()

    def refill(self):
        if len(self.d) - self.p >= 8:
            w = int.from_bytes(self.d[self.p:self.p + 7], "big")
            self.p += 7
            self.value = ((self.value << 56) | w) & ((1 << 64) - 1)
            self.bits += 56
        elif self.p < len(self.d):
            self.value = ((self.value << 8) | self.d[self.p]) & ((1 << 64) - 1)
            self.p += 1
            self.bits += 8
        elif not self.eof:
            self.value = (self.value << 8) & ((1 << 64) - 1)
            self.bits += 8
            self.eof = True
        else:
            self.bits = 0

    def get(self, prob):
        r = self.range
        if self.bits < 0:
            self.refill()
        pos = self.bits
        split = (r * prob) >> 8
        v = self.value >> pos
        if v > split:
            r -= split
            self.value -= (split + 1) << pos
            bit = 1
        else:
            r = split + 1
            bit = 0
        shift = 7 ^ (r.bit_length() - 1)
        r <<= shift
        self.bits -= shift
        self.range = r - 1
        return bit


class DevDec:  # wg::DBoolDec after adopt()
    def __init__(self, data, host):
        bits, value, p = host.bits, host.value, host.p
        if not host.eof:  # parse_frame: drop the whole bytes still unread in the window
            while bits >= 8:
                value >>= 8
                bits -= 8
                p -= 1
        while bits > 15:  # adopt(): EOF already hit with zero bits pending
            value >>= 8
            bits -= 8
        self.d, self.p, self.value, self.range, self.bits, self.eof = data, p, value & 0xffffffff, host.range, bits, host.eof

    def refill(self):
        if len(self.d) - self.p >= 2:
            self.value = ((self.value << 16) | (self.d[self.p] << 8) | self.d[self.p + 1]) & 0xffffffff
            self.p += 2
            self.bits += 16
        elif self.p < len(self.d):
            self.value = ((self.value << 8) | self.d[self.p]) & 0xffffffff
            self.p += 1
            self.bits += 8
        elif not self.eof:
            self.value = (self.value << 8) & 0xffffffff
            self.bits += 8
            self.eof = True
        else:
            self.bits = 0

    get = HostDec.get


def test_device_decoder_continues_the_host_decoder_bit_for_bit():
    rnd = random.Random(7)
    for trial in range(300):
        n = rnd.randint(0, 60)
        data = bytes(rnd.randint(0, 255) for _ in range(n))
        if data[:1] == b"\xff":
            # value >> bits <= range holds for every stream whose first byte is not 0xff (no encoder emits one: value < range + 1
            # = 255 at the start) and the narrow device window relies on it; the library parses such streams on the host
            data = b"\xfe" + data[1:]
        probs = [rnd.choice([1, 2, 17, 128, 145, 200, 254, 255, rnd.randint(1, 255)]) for _ in range(8 * n + 80)]
        ref = HostDec(data)
        expect = [(ref.get(p), ref.eof) for p in probs]
        for cut in sorted({0, 1, 2, 7, 8, 9, rnd.randint(0, len(probs)), rnd.randint(0, len(probs)), len(probs) // 2}):
            if cut > len(probs):
                continue
            host = HostDec(data)
            got = [(host.get(p), host.eof) for p in probs[:cut]]
            dev = DevDec(data, host)
            got += [(dev.get(p), dev.eof) for p in probs[cut:]]
            assert got == expect, (trial, n, cut)


def test_invariant_breaks_only_when_a_partition_starts_with_0xff():
    rnd = random.Random(9)
    for trial in range(200):
        data = bytes([rnd.randint(0, 254)] + [rnd.randint(0, 255) for _ in range(rnd.randint(0, 40))])
        d = HostDec(data)
        for _ in range(8 * len(data) + 40):
            d.get(rnd.randint(1, 255))
            if d.bits >= 0:
                assert (d.value >> d.bits) <= d.range
