"""The per-k-mer cache of scan_flags (faucet_b200/csrc/scan.cuh, scan_flags_memo_kernel) stores, instead of the
62/64-bit canonical k-mer, the home slot (implicit: slot index minus probe displacement) and the remaining bits
("quotient") of a 64-bit mix of it.  That identifies the k-mer EXACTLY only if the mix is a bijection on 64-bit
words.  This test restates the mix and inverts it: multiplication by an odd constant is invertible modulo 2^64 and
x ^= x >> 29 is undone by repeating the shift-xor; a cache hit therefore never belongs to another k-mer."""
import random

M64 = (1 << 64) - 1
C = 0x9E3779B97F4A7C15


def mix(x):
    h = (x * C) & M64
    return h ^ (h >> 29)


def unmix(h):
    x = h
    x ^= x >> 29
    x ^= x >> 58  # (h ^ h>>29) inverted: x = h ^ h>>29 ^ h>>58 ^ ...
    return (x * pow(C, -1, 1 << 64)) & M64


def test_mix_is_a_bijection():
    rng = random.Random(7)
    keys = [0, 1, M64 - 1, (1 << 62) - 1, 0x5555555555555555] + [rng.getrandbits(64) for _ in range(20000)]
    for k in keys:
        assert unmix(mix(k)) == k


def test_entry_identifies_the_kmer():
    """[0 | disp:3 | quotient:44 | masks:16]: (slot, entry) -> k-mer, for every table size the library uses"""
    rng = random.Random(11)
    for lg in (20, 24, 27, 30):
        qbits = 64 - lg
        assert qbits <= 44
        for _ in range(2000):
            key, disp, masks = rng.getrandbits(62), rng.randrange(8), rng.getrandbits(16)
            h = mix(key)
            home, quot = h >> qbits, h & ((1 << qbits) - 1)
            slot = (home + disp) & ((1 << lg) - 1)
            entry = (((disp << 44) | quot) << 16) | masks
            assert entry >> 63 == 0 and entry != M64          # never looks empty
            # decode
            d, q, m = (entry >> 60) & 7, (entry >> 16) & ((1 << 44) - 1), entry & 0xffff
            hm = (slot - d) & ((1 << lg) - 1)
            assert unmix((hm << qbits) | q) == key and m == masks
