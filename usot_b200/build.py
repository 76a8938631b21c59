"""In-tree nvcc build of libusot_b200.so (sm_100a only; no torch headers, plain C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libusot_b200.so")
SOURCES = ["kernels_simt.cu", "kernels_glue.cu", "conv_tc.cu", "stem_tc.cu", "xcorr_tma.cu", "pred_tma.cu", "crop.cu", "ops_abi.cu", "train_kernels.cu", "wgrad_tc.cu", "engine.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-shared", "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "usot_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into usot_b200/libusot_b200.so.  Returns the library path.
    Safe to call from several processes at once (torchrun ranks): an exclusive file lock serialises the build."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another rank built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    procs = []
    for s in srcs:  # compile translation units in parallel
        o = s[:-3] + ".o"
        objs.append(o)
        cmd = [_nvcc()] + [f for f in flags if f not in ("-shared",)] + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
