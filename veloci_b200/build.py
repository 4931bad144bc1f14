"""In-tree builds of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  lib/libveloci_b200.so        CUDA kernels + host runtime + C-ABI (nvcc, sm_100a)
  lib/libveloci_b200_index.so  index-building helpers for tests/bench (g++)
  oracle/_build/libveloci_oracle.so  CPU oracle, test infrastructure (g++)
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ROOT)
LIB_DIR = os.path.join(ROOT, "lib")
CSRC = os.path.join(ROOT, "csrc")

CXXFLAGS = ["-std=c++17", "-O3", "-march=x86-64-v3", "-fPIC", "-pthread", "-Wall", "-Wextra"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-pthread,-Wall,-march=x86-64-v3",
    "--expt-relaxed-constexpr", "-fmad=false",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)


def _headers():
    hs = []
    for pat in ("format/*.hpp", "host/*.hpp", "index/*.hpp", "cuda/*.cuh", "cuda/*.hpp", "*.hpp"):
        hs += glob.glob(os.path.join(CSRC, pat))
    hs += glob.glob(os.path.join(REPO, "include", "*.h"))
    return hs


def build_index_lib(force=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, "libveloci_b200_index.so")
    src = os.path.join(CSRC, "index", "index_capi.cpp")
    if force or _newer(out, [src] + _headers()):
        _run(["g++"] + CXXFLAGS + ["-shared", src, "-o", out])
    return out


def build_oracle(force=False):
    out_dir = os.path.join(REPO, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libveloci_oracle.so")
    src = os.path.join(REPO, "oracle", "veloci_oracle.cpp")
    if force or _newer(out, [src] + _headers()):
        _run(["g++"] + CXXFLAGS + ["-shared", src, "-o", out])
    return out


def build_main_lib(force=False, verbose_ptxas=False):
    """Every source is compiled to its own object (in parallel, only when it or a header changed) and linked."""
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    out = os.path.join(LIB_DIR, "libveloci_b200.so")
    srcs = sorted(glob.glob(os.path.join(CSRC, "cuda", "*.cu"))) + sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else [])
    if os.environ.get("VELOCI_PROBES"):  # timing experiments (tools/*_probe.py): the library reads its probe settings from the environment
        flags = flags + ["-DVELOCI_PROBES"]
    stamp = os.path.join(obj_dir, ".flags")
    flag_text = " ".join(flags)
    if not os.path.exists(stamp) or open(stamp).read() != flag_text:
        force = True
    headers = _headers()
    jobs = []
    objs = []
    for src in srcs:
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append([nvcc] + flags + ["-I", os.path.join(REPO, "include"), "-c", src, "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            list(pool.map(_run, jobs))
        open(stamp, "w").write(flag_text)
    if jobs or _newer(out, objs):
        _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", out, "-lcudart", "-ldl", "-lrt"])
    return out


def build_all(force=False):
    return [build_index_lib(force), build_oracle(force), build_main_lib(force)]


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
