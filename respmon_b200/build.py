"""Build librespmon_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The library is the product's only compute path; there is no CPU fallback.  `python -m respmon_b200.build`
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "librespmon_b200.so")
OBJ_DIR = os.path.join(PKG, "csrc", "_obj")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-O2", "--expt-relaxed-constexpr"] + os.environ.get("RM_NVCC_EXTRA", "").split()
# files whose double-precision scalar code mirrors SciPy/MINPACK operation by operation: no FMA contraction
NO_FMAD = {"signal.cu", "measure.cu"}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def build_variant(tag, extra_flags):
    """Developer builds for A/B timing: the same sources with extra nvcc flags (e.g. "-DPU_MAX_WARPS=24") linked into
    respmon_b200/_variants/librespmon_b200.<tag>.so.  Select one at run time with RESPMON_B200_LIB=<path> (_cabi.py);
    the product library is not touched."""
    nvcc = _nvcc()
    vdir = os.path.join(PKG, "_variants")
    odir = os.path.join(vdir, "_obj_" + tag)
    os.makedirs(odir, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    procs, objs = [], []
    for s in sources:
        obj = os.path.join(odir, s[:-3] + ".o")
        cmd = [nvcc, *ARCH, *COMMON, *extra_flags.split(), "-c", os.path.join(CSRC, s), "-o", obj]
        if s in NO_FMAD:
            cmd.insert(1, "-fmad=false")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s (%s):\n%s" % (s, tag, out))
    out_path = os.path.join(vdir, "librespmon_b200.%s.so" % tag)
    subprocess.check_call([nvcc, *ARCH, "-shared", "-o", out_path, *objs, "-Xcompiler", "-fPIC"])
    return out_path


def build(force=False, verbose=False):
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "respmon_b200.h"))
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp")
    digest = _digest([os.path.join(CSRC, s) for s in sources] + headers + [os.path.abspath(__file__)])
    digest += "|" + os.environ.get("RM_NVCC_EXTRA", "")     # developer flags (timing hooks) never masquerade as the product build
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in sources:
        obj = os.path.join(OBJ_DIR, s[:-3] + ".o")
        cmd = [nvcc, *ARCH, *COMMON, "-c", os.path.join(CSRC, s), "-o", obj]
        if s in NO_FMAD:
            cmd.insert(1, "-fmad=false")
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("[nvcc %s]\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, *ARCH, "-shared", "-o", OUT, *objs, "-Xcompiler", "-fPIC"])
    with open(stamp, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "--variant":      # python -m respmon_b200.build --variant w24 "-DPU_MAX_WARPS=24"
        print(build_variant(sys.argv[2], " ".join(sys.argv[3:])))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
