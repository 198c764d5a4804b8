"""Build the C-ABI shared library in-tree with nvcc for sm_100a.

    python -m neurecon_b200.build            # build if stale
    python -m neurecon_b200.build --force

The library lands in neurecon_b200/lib/libneurecon_b200.so (git-ignored, shipped to the GPU
box by gpurun).  There is no JIT and no fallback: if the library is missing the package
raises at first use.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libneurecon_b200.so")
# test-only twin of the two fused MLP kernels, compiled with the weight producer's fault injection (-DNR_FAULT_INJECT,
# csrc/umma.cuh); loaded by tests/test_gpu_reliability.py and tools/soak_mlp.py, never by the package
INJECT_LIB_PATH = os.path.join(LIB_DIR, "libneurecon_b200_inject.so")
INJECT_SOURCES = ("api.cu", "mlp_rev.cu", "mlp_rev_split.cu", "mlp_umma.cu")
# self-tests and micro-architecture probes (csrc/devtools/): their own library, the production one carries only the path
DEVTOOLS_LIB_PATH = os.path.join(LIB_DIR, "libneurecon_b200_devtools.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def devtools_sources():
    return sorted(glob.glob(os.path.join(CSRC, "devtools", "*.cu")))


def _stale():
    libs = (LIB_PATH, INJECT_LIB_PATH, DEVTOOLS_LIB_PATH)
    if not all(os.path.exists(p) for p in libs):
        return True
    t = min(os.path.getmtime(p) for p in libs)
    deps = sources() + devtools_sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile and link under an exclusive file lock: eight torchrun ranks that all find the library stale must not
    race nvcc into the same object files while their peers dlopen() the result."""
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    if not force and not _stale():     # re-checked under the lock: another rank may just have built it
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    objs, inject_objs = [], []
    procs = []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
        if os.path.basename(src) in INJECT_SOURCES:
            obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".inject.o")
            cmd = [nvcc, *NVCC_FLAGS, "-DNR_FAULT_INJECT", "-I", INCLUDE, "-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            inject_objs.append(obj)
    dev_objs = [os.path.join(LIB_DIR, "api.o")]
    for src in devtools_sources():
        obj = os.path.join(LIB_DIR, "devtools_" + os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        dev_objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("neurecon_b200: nvcc build failed")
    # link to a temporary name and rename: a process that dlopen()s the library meanwhile sees the old or the new file,
    # never a half-written one
    for path, members in ((LIB_PATH, objs), (INJECT_LIB_PATH, inject_objs), (DEVTOOLS_LIB_PATH, dev_objs)):
        tmp = "%s.%d.tmp" % (path, os.getpid())
        subprocess.check_call([nvcc, "-shared", "-o", tmp, *members, "-lcudart"])   # no -lcuda: see csrc/gemm16.cu
        os.replace(tmp, path)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
