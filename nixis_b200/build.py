"""Build libnixis_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m nixis_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libnixis_b200.so")
SOURCES = ["nxb_api.cu", "nxb_noise.cu", "nxb_mesh.cu", "nxb_adjacency.cu", "nxb_assembly.cu",
           "nxb_erosion.cu", "nxb_halo.cu", "nxb_noise_f64.cu", "nxb_climate.cu"]
# per-file extra flags: the reference-exact FP64 kernels must not contract a*b+c into FMA
EXTRA_FLAGS = {"nxb_noise_f64.cu": ["-fmad=false"], "nxb_climate.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--compiler-options", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v", "-DNXB_HAVE_NOISE4"] + (["-DNXB_ERO_PROFILE"] if os.environ.get("NXB_ERO_PROFILE") else [])


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "nixis_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    logs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + EXTRA_FLAGS.get(os.path.basename(s), []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            logs.append(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
            if verbose:
                sys.stderr.write(r.stderr)
    if force or _stale(SO, objs):
        cmd = [_nvcc(), "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    with open(os.path.join(objdir, "ptxas.log"), "a") as f:
        f.write("".join(logs))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
