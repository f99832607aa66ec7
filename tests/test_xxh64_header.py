"""omm_b200/csrc/omm_xxh64.h (the product's XXH64: byte-stream form for blob digests on the host, 32-bit-word streaming form for the LSH layer hashes on
the device) against the golden vectors of the vendored xxHash (tests/golden/xxh64.json, generated from external/xxHash by tests/golden/make_golden.py).
The header compiles for the host; a tiny C++ harness exposes both forms."""
import ctypes as C
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include "omm_xxh64.h"
extern "C" unsigned long long bytes_form(const void* p, size_t n, unsigned long long seed) { return ommb200::HostXxh64(p, n, seed); }
extern "C" unsigned long long words_form(const unsigned* w, size_t n, unsigned long long seed) {
    ommb200::xxh::WordStream s(seed);
    for (size_t i = 0; i < n; ++i) s.push(w[i]);
    return s.finish();
}
// the accumulator chain as the big-block digest kernel runs it (omm_bake.cu, ItemPostBigPipelined): pre-multiplied stripes, ChainStart /
// ChainStep / ChainEnd; n is a multiple of 32
extern "C" unsigned long long chain_form(const unsigned char* p, size_t n, unsigned long long seed) {
    using namespace ommb200::xxh;
    const unsigned long long acc0[4] = {seed + kP1 + kP2, seed + kP2, seed, seed - kP1};
    unsigned long long v[4];
    for (int j = 0; j < 4; ++j) {
        unsigned long long s = ChainStart(acc0[j]);
        for (size_t st = 0; st < n / 32; ++st) {
            unsigned long long in;
            memcpy(&in, p + 32 * st + 8 * j, 8);
            s = ChainStep(s, in * kP2);
        }
        v[j] = ChainEnd(s);
    }
    unsigned long long h = Rotl(v[0], 1) + Rotl(v[1], 7) + Rotl(v[2], 12) + Rotl(v[3], 18);
    for (int j = 0; j < 4; ++j) h = MergeRound(h, v[j]);
    return Avalanche(h + n);
}
extern "C" int chain_step_matches(unsigned long long s, unsigned long long x) {
    using namespace ommb200::xxh;
    return ChainStep(s, x) == Rotl(s, 31) * kP1 + x;
}
'''


def _xorshift_bytes(n):
    x, out = 88172645463325252, bytearray()
    for _ in range(n):
        x ^= (x << 13) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 7
        x ^= (x << 17) & 0xFFFFFFFFFFFFFFFF
        out.append((x >> 32) & 0xFF)
    return bytes(out)


def test_both_forms_match_the_vendored_xxhash():
    with tempfile.TemporaryDirectory() as d:
        src, so = os.path.join(d, "h.cpp"), os.path.join(d, "h.so")
        with open(src, "w") as f:
            f.write(HARNESS)
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "omm_b200", "csrc"), src, "-o", so])
        lib = C.CDLL(so)
        lib.bytes_form.restype, lib.bytes_form.argtypes = C.c_uint64, [C.c_char_p, C.c_size_t, C.c_uint64]
        lib.words_form.restype, lib.words_form.argtypes = C.c_uint64, [C.c_char_p, C.c_size_t, C.c_uint64]
        with open(os.path.join(ROOT, "tests", "golden", "xxh64.json")) as f:
            g = json.load(f)
        buf = _xorshift_bytes(5000)
        checked_words = 0
        for length, seed, want in g["plain"]:
            assert lib.bytes_form(buf, length, seed) == int(want), (length, seed)
            if length % 4 == 0:   # the streaming form takes whole 32-bit words
                assert lib.words_form(buf, length // 4, seed) == int(want), (length, seed)
                checked_words += 1
        assert checked_words >= 6
        for lvl, want in g["states"]:
            n = 1 << (2 * lvl)
            st = bytes((3 if (b % 3) == 2 else (b % 3)) for b in buf[:n])
            assert lib.bytes_form(st, n, 42) == int(want), lvl


def test_chain_form_of_the_big_block_digest():
    """ChainStart / ChainStep / ChainEnd (the funnel-shift, hand-split form of `acc <- rotl(acc + in * P2, 31) * P1` that the two-warp digest
    kernel of blocks of level >= 9 runs) against the plain byte-stream XXH64, and the step against its 64-bit definition."""
    import random
    with tempfile.TemporaryDirectory() as d:
        src, so = os.path.join(d, "h.cpp"), os.path.join(d, "h.so")
        with open(src, "w") as f:
            f.write(HARNESS)
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "omm_b200", "csrc"), src, "-o", so])
        lib = C.CDLL(so)
        lib.bytes_form.restype, lib.bytes_form.argtypes = C.c_uint64, [C.c_char_p, C.c_size_t, C.c_uint64]
        lib.chain_form.restype, lib.chain_form.argtypes = C.c_uint64, [C.c_char_p, C.c_size_t, C.c_uint64]
        lib.chain_step_matches.restype, lib.chain_step_matches.argtypes = C.c_int, [C.c_uint64, C.c_uint64]
        rng = random.Random(20261017)
        edge = [0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0x100000000, 0x7FFFFFFFFFFFFFFF, 0x8000000000000000, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF00000000, 0x00000001FFFFFFFF]
        for s in edge:
            for x in edge:
                assert lib.chain_step_matches(s, x), (hex(s), hex(x))
        for _ in range(20000):
            assert lib.chain_step_matches(rng.getrandbits(64), rng.getrandbits(64))
        buf = _xorshift_bytes(1 << 18)                                  # 4^9 bytes: the smallest block the kernel takes
        states = bytes((3 if (b % 3) == 2 else (b % 3)) for b in buf)   # 3-state values, like the hashed blocks
        for data in (buf, states):
            for n in (32, 64, 4096, 1 << 14, 1 << 18):
                for seed in (42, 0, 0xFFFFFFFFFFFFFFFF):
                    assert lib.chain_form(data, n, seed) == lib.bytes_form(data, n, seed), (n, seed)

