"""TEST INFRASTRUCTURE: builds tests/cuda_emu/_build/libpnr_emu_test.so, the SIMT translation units of
pixel-nerf_b200/csrc compiled with g++ against tests/cuda_emu/cuda_runtime.h (see there).  Only tests import this."""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pixel-nerf_b200", "csrc")
OUT = os.path.join(HERE, "_build")
UNITS = ["pnr_api.cu", "pnr_stages.cu", "pnr_field_simt.cu", "pnr_field_bwd.cu", "pnr_mgpu.cu"]
LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^;{}()]*?>)?)\s*<<<(.+?)>>>\s*\((.*?)\);", re.S)


def split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


KDEF = re.compile(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(")
DDEF = re.compile(r"__device__\s+(?:__forceinline__\s+)?[\w:<>]+[\s\*&]+(\w+)\s*\(")


def _body(text, start):
    i = text.find("{", start)
    j = text.find(";", start)
    if i < 0 or (0 <= j < i):
        return None
    depth, k = 0, i
    while True:
        depth += text[k] == "{"
        depth -= text[k] == "}"
        k += 1
        if depth == 0:
            return text[i:k]


def warp_collective_helpers():
    """Names of __device__ functions in the shared headers that use warp collectives: a kernel calling one of them
    needs its threads run as warps even if its own body does not mention __shfl / __syncwarp."""
    names = set()
    for name in sorted(os.listdir(CSRC)):
        if not name.endswith(".cuh") or name == "pnr_tc_ptx.cuh":
            continue
        text = open(os.path.join(CSRC, name)).read()
        for m in DDEF.finditer(text):
            body = _body(text, m.end())
            if body and ("__shfl" in body or "__syncwarp" in body):
                names.add(m.group(1))
    return names


def classify(texts):
    """kernel name -> emu::Mode, from what the kernel body uses."""
    modes = {}
    helpers = warp_collective_helpers()
    for text in texts:
        for m in KDEF.finditer(text):
            i = text.find("{", m.end())
            j = text.find(";", m.end())
            if i < 0 or (0 <= j < i):
                continue                     # a declaration, not a definition
            depth, k = 0, i
            while True:
                depth += text[k] == "{"
                depth -= text[k] == "}"
                k += 1
                if depth == 0:
                    break
            body = text[i:k]
            if "__syncthreads" in body:
                modes[m.group(1)] = "emu::BLOCK"
            elif "__shfl" in body or "__syncwarp" in body or any(re.search(r"\b%s\s*\(" % h, body) for h in helpers):
                modes[m.group(1)] = "emu::WARP"
            else:
                modes[m.group(1)] = "emu::SEQ"
    return modes


def rewrite(text, modes):
    def sub(m):
        kern, cfg, args = m.group(1), split_top(m.group(2)), m.group(3)
        smem = cfg[2] if len(cfg) > 2 else "0"
        mode = modes[kern.split("<")[0].split("::")[-1]]
        return (f"emu::launch<{mode}>(emu::to_dim3({cfg[0]}), emu::to_dim3({cfg[1]}), (size_t)({smem}), "
                f"[=]() {{ {kern}({args}); }});")
    text = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];",
                  r"\1* \2 = reinterpret_cast<\1*>(emu::blk->dyn.data());", text)
    return LAUNCH.sub(sub, text)


def build_asan():
    """Same sources with -fsanitize=address -> tests/cuda_emu/_build/libpnr_emu_asan.so.  Use it as
         ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
         PNR_EMU_LIB=tests/cuda_emu/_build/libpnr_emu_asan.so python -m pytest tests/test_emu_kernels.py
    to get a compute-sanitizer-like memcheck of the CUDA sources on the CPU (out-of-bounds on any tensor or
    workspace slice is reported with the .cu line)."""
    os.makedirs(OUT, exist_ok=True)
    texts = {u: open(os.path.join(CSRC, u)).read() for u in UNITS}
    modes = classify(texts.values())
    srcs = []
    for u in UNITS:
        dst = os.path.join(OUT, u.replace(".cu", "_asan.cpp"))
        open(dst, "w").write(rewrite(texts[u], modes))
        srcs.append(dst)
    srcs.append(os.path.join(HERE, "emu_stubs.cpp"))
    lib = os.path.join(OUT, "libpnr_emu_asan.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-fno-omit-frame-pointer",
                    "-fsanitize=address", "-ffp-contract=off", "-w", "-I", HERE, "-I", CSRC, "-o", lib] + srcs,
                   check=True)
    return lib


def build():
    os.makedirs(OUT, exist_ok=True)
    srcs, h = [], hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + sorted(os.listdir(HERE)):
        p = os.path.join(CSRC if os.path.exists(os.path.join(CSRC, name)) else HERE, name)
        if os.path.isfile(p) and name.split(".")[-1] in ("cu", "cuh", "h", "cpp", "py"):
            h.update(open(p, "rb").read())
    h.update(open(os.path.join(ROOT, "include", "pnr.h"), "rb").read())
    lib = os.path.join(OUT, "libpnr_emu_test.so")
    stamp = os.path.join(OUT, "stamp")
    if os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return lib
    texts = {u: open(os.path.join(CSRC, u)).read() for u in UNITS}
    modes = classify(texts.values())
    for u in UNITS:
        dst = os.path.join(OUT, u.replace(".cu", "_emu.cpp"))
        open(dst, "w").write(rewrite(texts[u], modes))
        srcs.append(dst)
    srcs.append(os.path.join(HERE, "emu_stubs.cpp"))
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fno-omit-frame-pointer", "-ffp-contract=off", "-w",
           "-I", HERE, "-I", CSRC, "-o", lib] + srcs
    subprocess.run(cmd, check=True)
    open(stamp, "w").write(h.hexdigest())
    return lib


if __name__ == "__main__":
    import sys
    print(build_asan() if "--asan" in sys.argv else build())
