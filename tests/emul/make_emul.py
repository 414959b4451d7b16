"""TEST INFRASTRUCTURE — build the whole of pysparselp_b200/csrc for the CPU on top of tests/emul/shim.

The CUDA sources are copied into tests/emul/_build/full/ with every
    kernel<<<grid, block, smem, stream>>>(args)
rewritten to
    emul::launch((unsigned)(grid), (unsigned)(block), [=]() { kernel(args); })
and compiled with g++ against the shim headers (cuda_runtime.h, cub/cub.cuh, nccl.h).  The result,
libcpppd_emul.so, exports the same C ABI as the product library; tests bind it directly with ctypes.
The product package never looks for it.
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pysparselp_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build", "full")
LIB = os.path.join(HERE, "_build", "libcpppd_emul.so")


def _match_forward(text, start, open_ch, close_ch):
    depth = 0
    for i in range(start, len(text)):
        ch = text[i]
        if ch == open_ch:
            depth += 1
        elif ch == close_ch:
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced %s%s" % (open_ch, close_ch))


def _split_top_level(text):
    parts, depth, cur = [], 0, []
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def rewrite_launches(text):
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            break
        # kernel name (with optional template arguments) right before <<<
        i = k
        while i > 0 and text[i - 1].isspace():
            i -= 1
        if text[i - 1] == ">":
            depth, j = 0, i - 1
            while True:
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
            i = j
        j = i
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        name = text[j:k].strip()
        end_cfg = text.index(">>>", k)
        cfg = _split_top_level(text[k + 3: end_cfg])
        a = end_cfg + 3
        while text[a].isspace():
            a += 1
        assert text[a] == "(", "launch of %s is not followed by an argument list" % name
        b = _match_forward(text, a, "(", ")")
        args = text[a + 1: b]
        out.append(text[pos:j])
        out.append("emul::launch((unsigned)(%s), (unsigned)(%s), [=]() { %s(%s); })" % (cfg[0], cfg[1], name, args))
        pos = b + 1
    return "".join(out)


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def build(force=False, defines=()):
    srcs = [os.path.join(CSRC, f) for f in sources()] + [os.path.join(ROOT, "include", "cpppd.h")] + [
        os.path.join(HERE, "shim", f) for f in ("cuda_runtime.h", "nccl.h", os.path.join("cub", "cub.cuh"))] + [
        os.path.join(HERE, "emul_nccl.cpp"), __file__]
    if not force and os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    header = os.path.join(ROOT, "include", "cpppd.h")
    for f in sources():
        text = open(os.path.join(CSRC, f)).read()
        text = text.replace('#include "../../include/cpppd.h"', '#include "%s"' % header)
        text = rewrite_launches(text)
        dst = f[:-3] + ".cpp" if f.endswith(".cu") else f
        with open(os.path.join(OUT_DIR, dst), "w") as fh:
            fh.write(text)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-w",
           "-I", os.path.join(HERE, "shim")] + list(defines) + os.environ.get("CPPPD_NVCC_DEFINES", "").split() + [
        "-o", LIB, os.path.join(OUT_DIR, "cpppd.cpp"), os.path.join(HERE, "emul_nccl.cpp"), "-ldl", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulated build failed:\n" + res.stderr[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
