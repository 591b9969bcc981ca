"""Build libibx.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libibx.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOSTFLAGS = "-fPIC,-fopenmp,-ffp-contract=off,-Wall,-Wno-unused-variable,-Wno-unused-function"
# The numerical kernels reproduce the reference's separate float32 roundings bit for bit: no FMA contraction.
# (A residual is a small difference of large fluxes; a 1-ulp change of a flux moves it by ~1e-3 relative.)
NOFMA = ["-fmad=false"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "--extended-lambda", "-Xcompiler", HOSTFLAGS]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(HERE, "..", "include", "ibx.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    jobs = []
    for src in _sources():
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src + ".o")
        if force or not os.path.exists(op) or os.path.getmtime(op) < max(os.path.getmtime(sp), hm):
            # march_fast.cu (option "arithmetic" = 1) is the one numerical TU compiled WITH contraction
            extra = NOFMA if src in ("ops.cu", "cfd.cu", "fused.cu", "tile.cu", "march.cu", "closures.cu", "gen.cu", "rans.cu") else []
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", op]
            if src.endswith(".cpp"):
                cmd = [NVCC] + COMMON + ["-x", "cu"] * 0 + ["-c", sp, "-o", op]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    if jobs:
        with ThreadPoolExecutor(max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
            for src, rc, out in ex.map(run, jobs):
                if verbose or rc != 0:
                    sys.stderr.write(f"[ibx build] {src}\n{out}\n")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ, s + ".o") for s in _sources()]
    if jobs or not os.path.exists(LIB) or force or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        # -z defs: an undefined symbol fails the build here instead of at dlopen on the GPU box
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-ldl", "-Xlinker", "-z", "-Xlinker", "defs"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libibx.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
