"""Builds tests/_build/libkernels_emu.so: the product's kernel sources compiled by g++ against the CPU emulation of
CUDA in tests/cuda_emu/cuda_runtime.h.  TEST ONLY.  The sources are used as they are except for two mechanical
rewrites g++ cannot parse around: `kernel<<<grid, block, smem, stream>>>(args);` becomes
`emu::launch(dim3(grid), dim3(block), [&]() { kernel(args); });` and `__shared__` becomes a function-local static
(`extern __shared__` an extern array defined by the harness)."""
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "bonej2_b200", "csrc")
_OUTDIR = os.path.join(_ROOT, "tests", "_build")
OUT = os.path.join(_OUTDIR, "libkernels_emu.so")
KERNEL_SOURCES = ["edt.cu", "ridge.cu", "paint_sep.cu", "paint_disc.cu", "finish.cu", "paint_scatter.cu", "phantom.cu"]
_DEPS = [os.path.join(_CSRC, f) for f in KERNEL_SOURCES + ["common.cuh", "paint_sep_core.cuh"]] + \
        [os.path.join(_HERE, f) for f in ("cuda_runtime.h", "harness.cpp", "emu_build.py")]


def _match_back(text, i, open_c, close_c):
    """text[i] == close_c: index of the matching open_c"""
    depth = 0
    while i >= 0:
        if text[i] == close_c:
            depth += 1
        elif text[i] == open_c:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced " + close_c)


def _match_fwd(text, i, open_c, close_c):
    depth = 0
    while i < len(text):
        if text[i] == open_c:
            depth += 1
        elif text[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced " + open_c)


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text):
    out, pos = "", 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            return out + text[pos:]
        # kernel expression: identifier, optionally with template arguments, right in front of <<<
        j = k - 1
        while text[j].isspace():
            j -= 1
        if text[j] == ">":
            j = _match_back(text, j, "<", ">") - 1
        while text[j].isalnum() or text[j] == "_" or text[j] == ":":
            j -= 1
        kernel = text[j + 1:k].strip()
        e = text.index(">>>", k)
        cfg = _split_top(text[k + 3:e])
        a0 = text.index("(", e)
        a1 = _match_fwd(text, a0, "(", ")")
        semi = text.index(";", a1)
        assert text[a1 + 1:semi].strip() == "", "unexpected text after a kernel launch"
        out += text[pos:j + 1] + "emu::launch(dim3(%s), dim3(%s), [&]() { %s%s; });" % (cfg[0], cfg[1], kernel, text[a0:a1 + 1])
        pos = semi + 1


def translate(name):
    text = open(os.path.join(_CSRC, name)).read()
    # device code that only exists under nvcc (TMA staging: `#if BJT_USE_TMA ... #endif`, off for host compilers) is
    # cut out before the launches are rewritten; the plain-load kernels beside it are what runs here
    text = re.sub(r"^#if BJT_USE_TMA\n.*?^#endif\n", "", text, flags=re.S | re.M)
    text = rewrite_launches(text)
    # the few inline-PTX statements: prefetches are no-ops; shared-window loads / stores of the painter's
    # active-set store go through the emulated window address (emu_shared_ptr in cuda_runtime.h)
    text = re.sub(r'asm volatile\("prefetch\.global\.L1[^"]*"[^;]*;', "/* prefetch */;", text)
    text, n1 = re.subn(r'asm volatile\("ld\.shared\.u64 %0, \[%1\];"\s*:\s*"=l"\((\w+)\)\s*:\s*"r"\((.*?)\)\);',
                       r"\1 = *emu_shared_ptr(\2);", text)
    text, n2 = re.subn(r'asm volatile\("st\.shared\.u64 \[%0\], %1;"\s*::\s*"r"\((.*?)\),\s*"l"\((\w+)\)\s*:\s*"memory"\);',
                       r"*emu_shared_ptr(\1) = \2;", text)
    assert "asm" not in re.sub(r"/\*.*?\*/|//[^\n]*", "", text, flags=re.S), name + ": inline asm the emulation does not know"
    text = re.sub(r"\bextern\s+__shared__\b", "extern", text)
    text = re.sub(r"\b__shared__\b", "static", text)
    return '#line 1 "%s"\n' % os.path.join(_CSRC, name) + text


# -fno-gnu-unique / -Bsymbolic: the two emulated libraries define the same symbols (kernels, the emulation's inline
# statics); each must bind to its own copies when both are loaded into one test process

class _BuildLock:
    """one builder at a time (pytest-xdist workers all find the library stale after a source change); the others wait
    and then find it fresh"""

    def __enter__(self):
        import fcntl
        os.makedirs(_OUTDIR, exist_ok=True)
        self.f = open(os.path.join(_OUTDIR, ".build.lock"), "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)

    def __exit__(self, *a):
        self.f.close()


def build(force=False):
    with _BuildLock():
        return _build(force)


def _build(force):
    stale = (not os.path.exists(OUT)) or any(os.path.getmtime(p) > os.path.getmtime(OUT) for p in _DEPS)
    if not (force or stale):
        return OUT
    os.makedirs(_OUTDIR, exist_ok=True)
    objs = []
    for name in KERNEL_SOURCES:
        cpp = os.path.join(_OUTDIR, "emu_" + name.replace(".cu", ".cpp"))
        with open(cpp, "w") as f:
            f.write(translate(name))
        objs.append(cpp)
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-fno-gnu-unique", "-Wl,-Bsymbolic", "-I", _HERE, "-I", _CSRC, "-I", os.path.join(_ROOT, "include"),
           "-o", OUT] + objs + [os.path.join(_HERE, "harness.cpp")]
    subprocess.check_call(cmd[:cmd.index("-o") + 1] + [OUT + ".tmp"] + cmd[cmd.index("-o") + 2:])
    os.replace(OUT + ".tmp", OUT)
    return OUT


API_OUT = os.path.join(_OUTDIR, "libbonejthickness_emu.so")
API_SOURCES = ["api.cu", "edt.cu", "ridge.cu", "paint_sep.cu", "paint_disc.cu", "paint_scatter.cu", "finish.cu", "phantom.cu", "segstats.cu",
               "ridgepoints.cu", "peer_copy.cu"]


def build_api(force=False):
    """the whole C-ABI (include/bonej_thickness.h, api.cu + thickness_wrapper.cpp) on the emulation, for tests that
    drive the library's entry points without a GPU"""
    with _BuildLock():
        return _build_api(force)


def _build_api(force):
    deps = [os.path.join(_CSRC, f) for f in API_SOURCES + ["common.cuh", "paint_sep_core.cuh", "thickness_wrapper.cpp"]] + \
           [os.path.join(_HERE, f) for f in ("cuda_runtime.h", "api_stubs.cpp", "emu_build.py")] + \
           [os.path.join(_ROOT, "include", f) for f in ("bonej_thickness.h", "thickness_wrapper.hpp")]
    stale = (not os.path.exists(API_OUT)) or any(os.path.getmtime(p) > os.path.getmtime(API_OUT) for p in deps)
    if not (force or stale):
        return API_OUT
    os.makedirs(_OUTDIR, exist_ok=True)
    srcs = []
    for name in API_SOURCES:
        cpp = os.path.join(_OUTDIR, "emuapi_" + name.replace(".cu", ".cpp"))
        with open(cpp, "w") as f:
            f.write(translate(name))
        srcs.append(cpp)
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-fno-gnu-unique", "-Wl,-Bsymbolic", "-I", _HERE, "-I", _CSRC, "-I", os.path.join(_ROOT, "include"),
           "-o", API_OUT] + srcs + [os.path.join(_CSRC, "thickness_wrapper.cpp"), os.path.join(_HERE, "api_stubs.cpp"), "-lpthread"]
    subprocess.check_call(cmd[:cmd.index("-o") + 1] + [API_OUT + ".tmp"] + cmd[cmd.index("-o") + 2:])
    os.replace(API_OUT + ".tmp", API_OUT)
    return API_OUT


if __name__ == "__main__":
    print(build_api(force=True))
    print(build(force=True))
