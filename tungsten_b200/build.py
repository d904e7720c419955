"""Builds tungsten_b200/libtgb200.so (hand-written sm_100a CUDA + host C++) with nvcc, in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtgb200.so")
SOURCES = ["tgb200_api.cu", "bvh_build.cpp", "hair_tables.cpp", "sobol_blob.cpp"]
HEADERS = ["tgb_device.cuh", "tgb_wavefront.cuh", "bvh_build.h", "hair_tables.h", os.path.join("..", "..", "include", "tgb200.h")]
BLOB = os.path.join(HERE, "data", "sobol_1024x32.u32")

# -fmad=false: the reference is built without FMA contraction (CMakeLists.txt:17-19); radiance parity needs
# the same rounding.  No --use_fast_math: IEEE division / sqrt.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "--prec-div=true", "--prec-sqrt=true", "-Xcompiler", "-fPIC,-O3,-ffp-contract=off",
              "-Xptxas", "-v", "--shared", '-DTGB_SOBOL_BLOB_PATH="%s"' % BLOB]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [BLOB, os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Experimental build with extra -D flags -> tungsten_b200/libtgb200_<name>.so (load with TGB200_LIB)."""
    out = os.path.join(HERE, "libtgb200_%s.so" % name)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out, "-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed")
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT, "-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed, see %s" % log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
