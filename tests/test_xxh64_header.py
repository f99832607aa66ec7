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
