"""In-tree build of the native libraries (nvcc cross-compiles sm_100a without a GPU).

    python -m compression_algorithms_b200.build

Produces, next to this file:
  libb200comp.so    CUDA kernels + the device-level C-ABI (include/b200comp.h)
  libb200corpus.so  seeded synthetic corpus generator (host only)
  libb200_{huffman,lz77,deflate,fse}.so  drop-in shims exporting the reference's own
                    function names (one per reference directory, symbols collide)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--threads", "4"]


def _gcc():
    return "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def _gxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + cmd[0])
    return r.stdout


def build_core(force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cu = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = cu + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "b200comp.h")]
    out = os.path.join(HERE, "libb200comp.so")
    if force or _newer(out, deps):
        env_cc = ["-ccbin", _gxx()]
        _run([nvcc] + NVCC_FLAGS + env_cc + ["-shared"] + cu + ["-o", out])
    return out


def build_corpus(force=False):
    src = os.path.join(CSRC, "corpus.c")
    out = os.path.join(HERE, "libb200corpus.so")
    if force or _newer(out, [src]):
        _run([_gcc(), "-O3", "-march=x86-64-v3", "-fPIC", "-fopenmp", "-shared", src, "-o", out])
    return out


def build_shims(force=False):
    outs = []
    shim_dir = os.path.join(CSRC, "shims")
    if not os.path.isdir(shim_dir):
        return outs
    core = os.path.join(HERE, "libb200comp.so")
    for name in ("huffman", "lz77", "deflate", "fse"):
        src = os.path.join(shim_dir, "shim_%s.c" % name)
        if not os.path.exists(src):
            continue
        out = os.path.join(HERE, "libb200_%s.so" % name)
        deps = [src, core] + [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
        if force or _newer(out, deps):
            _run([_gcc(), "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                  "-L", HERE, "-lb200comp", "-Wl,-rpath,$ORIGIN", "-Wl,-Bsymbolic"])
        outs.append(out)
    return outs


def build_all(force=False):
    return [build_core(force), build_corpus(force)] + build_shims(force)


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv):
        print(p)
