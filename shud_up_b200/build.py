"""Build the CUDA library in-tree: shud_up_b200/libshud_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libshud_b200.so")
SOURCES = ["shud_rhs.cu", "shud_nvec.cu", "shud_io.cu", "shud_nvector_sundials.cu", "shud_nvector_generic.cpp",
           "shud_cvode.cpp"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              # the x86-64 reference build contracts no product-sums; neither do we (parity first)
              "-fmad=false",
              # static divisors (area, Sy, Dist2Nabor, avgRough) are stored as reciprocals: x*(1/d) instead of
              # x/d, <= 1.5 ulp apart, inside the 1e-12 parity tolerance (shud_phys.cuh)
              "-DSHUD_RCP",
              "-Xcompiler", "-fPIC", "-shared", "-ldl", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", h) for h in os.listdir(os.path.join(ROOT, "include"))] + [
                                                                 os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, extra=()):
    """out/extra: experimental variant builds (e.g. extra=["-DSHUD_X=1"], out="/path/lib_x.so")"""
    if out is None and not force and not needs_build():
        return LIB
    out = out or LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    flags = [f for f in NVCC_FLAGS if not (f == "-fmad=false" and any(e.startswith("-fmad") for e in extra))]
    cmd = [nvcc] + flags + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
