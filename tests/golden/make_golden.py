#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the reference checkout (run in the build container, where /root/reference exists):

  texcoord_kat.json     every TexCoordTest(...) expectation of support/tests/test_texture.cpp (address-mode KATs)
  xxh64.json            XXH64(data, len, seed) of the vendored external/xxHash for lengths that hit every code path
  std_hash_float.json   libstdc++ std::hash<float> (what glm's std::hash<vec2> -> the SDK's UV pre-dedup key is built from)
  morton.json           xy_to_morton of src/util/bit_tricks.h
  bake_digests.json     sha256 of the five result arrays of the UNMODIFIED SDK build (oracle/_ref) for a set of parity cases
  serialized_inputs.json  the DeserializeInput_* golden blobs (SDK 1.4 - 1.7 formats, plain and LZ4-compressed) of
                        support/tests/test_omm_bake_cpu.cpp with the state totals the SDK's tests expect after baking them

Usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODES = {"Wrap": 0, "Mirror": 1, "Clamp": 2, "Border": 3, "MirrorOnce": 4}


def texcoord_kats():
    src = open(os.path.join(REF, "support/tests/test_texture.cpp")).read()
    pat = re.compile(r"TexCoordTest\(omm::TextureAddressMode::(\w+),\s*\{\s*(-?\d+),\s*(-?\d+)\s*\},\s*\{\s*(-?\d+),\s*(-?\d+)\s*\},\s*([^;]+)\);")
    border = 0x7FFFFFFE

    def val(tok):
        tok = tok.strip()
        return border if "kTexCoordBorder" in tok else int(tok)

    out = []
    for m in pat.finditer(src):
        mode, x, y, w, h, exp = m.groups()
        exp = exp.strip()
        if "kTexCoordBorder2" in exp:
            ex, ey = border, border
        else:
            inner = exp[exp.index("{") + 1:exp.rindex("}")]
            toks = [t for t in inner.split(",") if t.strip()]
            if len(toks) != 2:
                continue
            ex, ey = val(toks[0]), val(toks[1])
        out.append([MODES[mode], int(x), int(y), int(w), int(h), ex, ey])
    return out


def serialized_inputs():
    src = open(os.path.join(REF, "support/tests/test_omm_bake_cpu.cpp")).read()
    out = []
    for m in re.finditer(r"TEST_P\(OMMBakeTestCPU, (DeserializeInput\w+)\)\s*\{(.*?)\n\t\}", src, re.S):
        name, body = m.group(1), m.group(2)
        arr = re.search(r"=\s*\{(.*?)\};", body, re.S)
        if not arr:
            continue
        data = bytes(int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]{2})", arr.group(1)))
        exp = {k: int(v) for k, v in re.findall(r"\.(\w+)\s*=\s*(\d+)", body)}
        out.append({"name": name, "line": src[:m.start()].count("\n") + 1, "blob_hex": data.hex(), "expect": exp})
    return {"source": "support/tests/test_omm_bake_cpu.cpp (golden serialized inputs of SDK versions 1.4 - 1.7 with the expected ommDebugGetStats totals)",
            "cases": out}


def run_c(code, lang, extra):
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "g." + ("cpp" if lang == "c++" else "c"))
        exe = os.path.join(td, "g")
        open(src, "w").write(code)
        cc = "/usr/bin/g++" if lang == "c++" else "/usr/bin/gcc"
        subprocess.check_call([cc, "-O1", "-o", exe, src] + extra)
        return subprocess.check_output([exe]).decode()


def xxh64_vectors():
    code = r'''
#include <stdio.h>
#include <stdint.h>
#include "xxhash.h"
int main(){ static uint8_t buf[5000]; uint64_t s=88172645463325252ull; for(int i=0;i<5000;i++){ s^=s<<13; s^=s>>7; s^=s<<17; buf[i]=(uint8_t)(s>>32);} 
 int lens[]={0,1,3,4,5,8,11,16,17,31,32,33,63,64,100,256,1024,4096,4999};
 for(unsigned k=0;k<sizeof(lens)/sizeof(int);k++) for(int seed=0; seed<2; seed++) printf("%d %d %llu\n", lens[k], seed?42:0, (unsigned long long)XXH64(buf,lens[k],seed?42:0));
 /* the shapes the baker hashes: 4^level bytes of values {0,1,3} */
 for(int lvl=0; lvl<=6; lvl++){ int n=1<<(2*lvl); static uint8_t st[4096]; for(int i=0;i<n;i++){ int v=buf[i]%3; st[i]=v==2?3:v;} printf("S %d %llu\n", lvl, (unsigned long long)XXH64(st,n,42)); }
 return 0; }
'''
    out = run_c(code, "c", ["-I" + os.path.join(REF, "external/xxHash"), os.path.join(REF, "external/xxHash/xxhash.c")])
    plain, states = [], []
    for ln in out.splitlines():
        f = ln.split()
        if f[0] == "S":
            states.append([int(f[1]), f[2]])
        else:
            plain.append([int(f[0]), int(f[1]), f[2]])
    return {"generator": "xorshift64 seed 88172645463325252, byte = state>>32", "plain": plain, "states": states}


def std_hash_float_vectors():
    code = r'''
#include <cstdio>
#include <functional>
#include <cstring>
#include <cstdint>
int main(){ const uint32_t bits[]={0u,0x80000000u,0x3f800000u,0xbf800000u,0x3f000000u,0x3e99999au,0x00000001u,0x7f7fffffu,0x3eaaaaabu,0x41200000u,0x3dcccccdu,0x40490fdbu};
 for(unsigned i=0;i<sizeof(bits)/4;i++){ float f; memcpy(&f,&bits[i],4); printf("%u %llu\n", bits[i], (unsigned long long)std::hash<float>()(f)); } return 0; }
'''
    out = run_c(code, "c++", [])
    return [[int(a), b] for a, b in (ln.split() for ln in out.splitlines())]


def morton_vectors():
    pts = [(0, 0), (1, 0), (0, 1), (3, 5), (255, 255), (1023, 1), (8191, 8191), (4096, 123), (65535, 65535)]
    code = "#include <cstdio>\n#include \"util/bit_tricks.h\"\nint main(){\n" + "".join(
        f'printf("%u\\n", omm::xy_to_morton({x}u,{y}u));\n' for x, y in pts) + "return 0;}\n"
    out = run_c(code, "c++", ["-std=gnu++20", "-I" + os.path.join(REF, "libraries/omm-lib/src"), "-I" + os.path.join(REF, "libraries/omm-lib/include"),
                              "-I" + os.path.join(REF, "external/glm")])
    return [[x, y, int(v)] for (x, y), v in zip(pts, out.split())]


DIGEST_CASES = ["c1_quad_checker_l3_2state", "c2_small_l4", "c3_small_l5", "c3_small_sat", "c5_small_mixed_levels", "addr_mirror_npot",
                "promo_nearest_4state", "filter_nearest_mips", "mips_linear", "degenerate_and_nan", "per_triangle_levels", "reuse_uv_and_content",
                "states_le_unknown_opaque", "dynamic_levels_area", "degenerate_dynamic_levels"]


def result_digests(res):
    return {k: hashlib.sha256(getattr(res, k).tobytes()).hexdigest() for k in ("array_data", "desc_array", "desc_histogram", "index_buffer", "index_histogram")} | {
        "index_format": int(res.index_format), "array_bytes": int(res.array_data.size), "descs": int(res.desc_array.size)}


def bake_digests():
    import parity_cases as PC
    from omm_b200.capi import OmmLib
    ref = OmmLib(os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so"))
    cases = PC.cases()
    out = {}
    for name in DIGEST_CASES:
        mk, ov = cases[name]
        out[name] = result_digests(PC.run_bake(ref, mk(), **ov))
    for name, (mk, ov) in PC.sdk_only_cases().items():   # near-duplicate merge / Compress: the port does not restate them
        out["sdk_only:" + name] = result_digests(PC.run_bake(ref, mk(), **ov))
    return out


def main():
    def dump(name, obj):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(obj, f, indent=0, separators=(",", ":"))
            f.write("\n")
        print("wrote", name)
    dump("texcoord_kat.json", texcoord_kats())
    dump("xxh64.json", xxh64_vectors())
    dump("std_hash_float.json", std_hash_float_vectors())
    dump("morton.json", morton_vectors())
    dump("bake_digests.json", bake_digests())
    dump("serialized_inputs.json", serialized_inputs())


if __name__ == "__main__":
    main()
