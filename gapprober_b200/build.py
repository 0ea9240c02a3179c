"""In-tree build of the native parts: libfigbird_cuda.so (sm_100a kernels + C ABI) and the host tools.

Explicit nvcc/g++ command lines (no JIT cache) so the built files sit next to the sources and travel
with the repository snapshot.  nvcc cross-compiles sm_100a without a GPU.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
TOOLS = os.path.join(HERE, "tools")
BIN = os.path.join(HERE, "bin")
LIB = os.path.join(HERE, "libfigbird_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # The reference's arithmetic has no fused multiply-add (x86-64 baseline); contraction would change bits.
    "-fmad=false",
    "-Xcompiler", "-fPIC,-Wall",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("fb_kernels.cu", "fb_microbench.cu", "fb_context.cpp")]
    deps = srcs + [os.path.join(CSRC, "fb_device.h"), os.path.join(ROOT, "include", "figbird_cuda.h")]
    if force or _newer(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs
        subprocess.check_call(cmd)
    return LIB


def build_tools(force=False):
    os.makedirs(BIN, exist_ok=True)
    out = []
    host = os.path.join(CSRC, "host")
    for name in ("fb_synth", "fb_split"):
        src = os.path.join(TOOLS, name + ".cpp")
        exe = os.path.join(BIN, name)
        deps = [src] + ([os.path.join(host, "fb_split.hpp"), os.path.join(host, "fb_io.hpp")] if name == "fb_split" else [])
        if force or _newer(exe, deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, src])
        out.append(exe)
    return out


def build_host(force=False):
    """The drop-in figbird binary (host preamble + file contract; dlopens libfigbird_cuda.so at run time)."""
    os.makedirs(BIN, exist_ok=True)
    host = os.path.join(CSRC, "host")
    src = os.path.join(host, "figbird.cpp")
    deps = [src, os.path.join(host, "fb_preamble.hpp"), os.path.join(host, "fb_io.hpp"), os.path.join(host, "fb_split.hpp"),
            os.path.join(ROOT, "include", "figbird_cuda.h")]
    exe = os.path.join(BIN, "figbird.out")
    if force or _newer(exe, deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
    return exe


def learn_model(workdir, scaffolds="scaffolds.fa", out="fb_model.bin"):
    """Run the host preamble of the drop-in binary in a GapProber working directory; returns the model path."""
    exe = build_host()
    env = dict(os.environ, FIGBIRD_MODEL_OUT=out, FIGBIRD_PREAMBLE_ONLY="1")
    subprocess.run([exe, scaffolds, " 0", "250"], cwd=workdir, env=env, check=True, stdout=subprocess.DEVNULL)
    return os.path.join(workdir, out)


def build_all(force=False, verbose=False):
    build_library(force, verbose)
    build_tools(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force=True, verbose=True)
