"""The arithmetic of k_pack_slices' single-pass CRC (csrc/ffv1_enc_kernels.cu), restated in Python and checked against the
plain byte-wise CRC (libavutil AV_CRC_32_IEEE as the oracle computes it, crc.c:357-380): thread t of 256 keeps the
remainder of the words t, t + 256, ... by Horner's rule with the constant multiplier x^8192 (four table lookups on the
bytes of the running remainder), scales it to the end of the message and the XOR of all threads is the slice CRC.  Runs
on the CPU; the GPU tests compare whole packets (which end in this CRC) with the reference's."""
import random
import pytest
from oracle import ffv1_oracle as O

P, M, T = 0x04C11DB7, 0xFFFFFFFF, 256

def gf_mulmod(a, b):                     # a(x) * b(x) mod P(x), bit k = coefficient of x^k
    r = 0
    for i in range(31, -1, -1):
        r = ((r << 1) & M) ^ (P if r & 0x80000000 else 0)
        if (b >> i) & 1:
            r ^= a
    return r

TAB0 = []
for n in range(256):
    c = n << 24
    for _ in range(8):
        c = ((c << 1) & M) ^ P if c & 0x80000000 else (c << 1) & M
    TAB0.append(c)
POW = []
p = 0x100
for _ in range(32):
    POW.append(p)
    p = gf_mulmod(p, p)
MUL = [[gf_mulmod((b << (8 * k)) & M, POW[10]) for b in range(256)] for k in range(4)]     # x^(32 * 256) = x^(8 * 2^10)

def bytewise(data):
    c = 0
    for b in data:
        c = ((c << 8) & M) ^ TAB0[(c >> 24) ^ b]
    return c

def scale(c, nbytes):
    m, j = 1, 0
    while nbytes:
        if nbytes & 1:
            m = gf_mulmod(m, POW[j])
        nbytes >>= 1
        j += 1
    return gf_mulmod(c, m)

def kernel_crc(payload, head, has_len=True):
    nb = len(payload)
    head = min(nb, head)
    nwords = (nb - head) >> 2
    done = head + 4 * nwords
    tail = ([(nb >> 16) & 255, (nb >> 8) & 255, nb & 255] if has_len else []) + [0]
    total = 0
    for t in range(T):
        c, last = 0, None
        for w in range(t, nwords, T):
            c = MUL[0][c & 255] ^ MUL[1][(c >> 8) & 255] ^ MUL[2][(c >> 16) & 255] ^ MUL[3][c >> 24]
            c ^= int.from_bytes(payload[head + 4 * w:head + 4 * w + 4], "big")
            last = w
        if last is not None and c:
            total ^= scale(c, (nwords - 1 - last) * 4 + (nb - done) + len(tail) + 4)
    h = bytewise(payload[:head])
    if h:
        total ^= scale(h, nb - head + len(tail))
    return total ^ bytewise(list(payload[done:]) + tail), bytes(payload) + bytes(tail)

@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 8, 100, 1023, 1024, 1025, 1028, 4099, 20011])
def test_strided_horner_crc_equals_the_bytewise_crc(n):
    rng = random.Random(n)
    for head in range(4):
        for has_len in (True, False):
            payload = bytes(rng.randrange(256) for _ in range(n))
            got, message = kernel_crc(payload, head, has_len)
            assert got == bytewise(message) == O.crc32(message), (n, head, has_len)
