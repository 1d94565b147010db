"""Builds tests/simt_emu/_build/libses_simt_emu.so: the engine's CUDA sources (simple-es_b200/csrc) compiled for the
HOST on top of the SIMT emulator in include/simt_emu.h.  TEST INFRASTRUCTURE ONLY -- see that header; the product
never loads this library.

The sources are used as they are; three textual rewrites make them C++:
  * kernel<<<grid, block[, smem[, stream]]>>>(args);   ->  simt::launch(grid, block, smem, [&]() { kernel(args); });
  * extern __shared__ ... unsigned char name[];        ->  unsigned char *name = simt::dyn_smem();
  * the inline-PTX statements (rcp.approx f32 / f64, min.xorsign.abs, %lanemask_lt, st.release / ld.acquire) -> C++ equivalents
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "simple-es_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libses_simt_emu.so")
# -DSES_BUILD_TESTS: the emulated library is test infrastructure -- it carries the test hooks and every alternative kernel
CXXFLAGS = ["-DSES_BUILD_TESTS", "-std=c++17", "-O2", "-g", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-fno-strict-aliasing",
            "-Wno-unknown-pragmas", "-Wno-attributes", "-Wno-unused-value"]


def _match_back(s, i, open_c, close_c):
    """s[i] == close_c: index of the matching open_c."""
    depth = 0
    while i >= 0:
        if s[i] == close_c:
            depth += 1
        elif s[i] == open_c:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced")


def _match_fwd(s, i, open_c, close_c):
    depth = 0
    while i < len(s):
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src):
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            return out + src[pos:]
        # kernel expression: identifier, optionally followed by <template arguments>
        j = i - 1
        while src[j].isspace():
            j -= 1
        if src[j] == ">":
            j = _match_back(src, j, "<", ">") - 1
        while j >= 0 and (src[j].isalnum() or src[j] in "_:"):
            j -= 1
        kernel = src[j + 1:i].strip()
        k = src.index(">>>", i)
        cfg = _split_top(src[i + 3:k])
        a0 = src.index("(", k)
        a1 = _match_fwd(src, a0, "(", ")")
        semi = src.index(";", a1)
        assert src[a1 + 1:semi].strip() == "", src[a1:semi + 1]
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[pos:j + 1] + "simt::launch(%s, %s, %s, [&]() { %s%s; });" % (grid, block, smem, kernel, src[a0:a1 + 1])
        pos = semi + 1


ASM = [
    (re.compile(r'asm\("rcp\.approx\.ftz\.f64 %0, %1;"\s*:\s*"=d"\((.+?)\)\s*:\s*"d"\((.+?)\)\);'), r"\1 = simt::rcp_approx_f64(\2);"),
    (re.compile(r'asm\("rcp\.approx\.ftz\.f32 %0, %1;"\s*:\s*"=f"\((.+?)\)\s*:\s*"f"\((.+?)\)\);'), r"\1 = simt::rcp_approx(\2);"),
    (re.compile(r'asm\("min\.xorsign\.abs\.f32 %0, %1, %2;"\s*:\s*"=f"\((.+?)\)\s*:\s*"f"\((.+?)\),\s*"f"\((.+?)\)\);'),
     r"\1 = simt::min_xorsign_abs(\2, \3);"),
    (re.compile(r'asm volatile\("mov\.u32 %0, %%lanemask_lt;"\s*:\s*"=r"\((.+?)\)\);'), r"\1 = simt::lanemask_lt();"),
    (re.compile(r'asm volatile\("mov\.u32 %0, %%smid;"\s*:\s*"=r"\((.+?)\)\);'), r"\1 = simt::smid();"),
    (re.compile(r'asm volatile\("bar\.sync %0, (\d+);"\s*::\s*"r"\((.+?)\)\s*:\s*"memory"\);'), r"simt::named_barrier(\2, \1);"),
    (re.compile(r'asm volatile\("st\.release\.sys\.global\.u64 \[%0\], %1;"\s*::\s*"l"\((.+?)\),\s*"l"\((.+?)\)\s*:\s*"memory"\);'),
     r"__atomic_store_n((unsigned long long *)(\1), (unsigned long long)(\2), __ATOMIC_RELEASE);"),
    (re.compile(r'asm volatile\("ld\.acquire\.sys\.global\.u64 %0, \[%1\];"\s*:\s*"=l"\((.+?)\)\s*:\s*"l"\((.+?)\)\s*:\s*"memory"\);'),
     r"\1 = __atomic_load_n((unsigned long long *)(\2), __ATOMIC_ACQUIRE);"),
]
DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char (\w+)\[\];")


def transform(text):
    text = rewrite_launches(text)
    text = DYN_SMEM.sub(r"unsigned char *\1 = simt::dyn_smem();", text)
    for rx, rep in ASM:
        text = rx.sub(rep, text)
    text = text.replace('#include "../../include/ses_b200.h"', '#include "%s"' % os.path.join(ROOT, "include", "ses_b200.h"))
    text = text.replace('#include "../../include/ses_b200_test.h"', '#include "%s"' % os.path.join(ROOT, "include", "ses_b200_test.h"))
    if re.search(r"\basm\b", text) or "<<<" in text:
        raise RuntimeError("simt_emu: an inline-asm statement or a kernel launch was not rewritten")
    return text


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def build(force=False):
    if os.environ.get("SES_SIMT_EMU_LIB"):             # e.g. a sanitizer build of the same sources (tools/emu_ubsan.sh)
        return os.environ["SES_SIMT_EMU_LIB"]
    deps = sources() + [os.path.join(HERE, "include", "simt_emu.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    src_dir = os.path.join(BUILD, "src")
    os.makedirs(src_dir, exist_ok=True)
    for path in sources():
        with open(path) as f:
            text = transform(f.read())
        name = os.path.basename(path)
        if name.endswith(".cu"):
            name = name[:-3] + ".cpp"
        with open(os.path.join(src_dir, name), "w") as f:
            f.write(text)
    cmd = ["g++"] + CXXFLAGS + ["-I", os.path.join(HERE, "include"), "-o", LIB, os.path.join(src_dir, "ses_abi.cpp")]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
