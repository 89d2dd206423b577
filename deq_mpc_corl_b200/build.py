"""Builds libalmpc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m deq_mpc_corl_b200.build [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "almpc_kernels.cu")
OUT = os.path.join(HERE, "libalmpc.so")
DEPS = [SRC, os.path.join(ROOT, "include", "almpc.h")] + [
    os.path.join(HERE, "csrc", h) for h in ("almpc_models.cuh", "almpc_common.cuh", "almpc_dynamics_kernels.cuh",
                                            "almpc_factor.cuh", "almpc_factor_mma.cuh", "almpc_linesearch.cuh", "almpc_fused.cuh", "almpc_kkt.cuh")]


def nvcc_path():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    """Four translation units ((fp64 | fp32) x (entry points | state-estimator variant)) compiled in parallel,
    then linked."""
    if not force and not needs_build():
        return OUT
    import concurrent.futures

    common = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc")]
    if verbose:
        common += ["-Xptxas", "-v"]
    common += os.environ.get("ALMPC_NVCC_FLAGS", "").split()  # experiment builds (-D switches)
    out = os.environ.get("ALMPC_OUT", OUT)
    objdir = os.path.join(HERE, "build" + os.environ.get("ALMPC_OBJ_SUFFIX", ""))
    os.makedirs(objdir, exist_ok=True)
    units = [("f64", ["-DALMPC_BUILD_F64=1", "-DALMPC_BUILD_F32=0"]), ("f32", ["-DALMPC_BUILD_F64=0", "-DALMPC_BUILD_F32=1"]),
             ("f64_se", ["-DALMPC_BUILD_F64=1", "-DALMPC_BUILD_F32=0", "-DALMPC_BUILD_SE=1"]),
             ("f32_se", ["-DALMPC_BUILD_F64=0", "-DALMPC_BUILD_F32=1", "-DALMPC_BUILD_SE=1"])]

    def compile_unit(u):
        name, defs = u
        obj = os.path.join(objdir, f"almpc_{name}.o")
        cmd = common + defs + ["-c", SRC, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:
        results = list(ex.map(compile_unit, units))
    if verbose:
        for _, err in results:
            print(err)
    tmp_out = out + ".tmp"  # link next to the target, then rename: a reader never sees a half-written library
    link = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp_out] + [o for o, _ in results]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
    os.replace(tmp_out, out)
    for o, _ in results:  # objects are not needed once linked (keeps the gpurun snapshot small)
        os.remove(o)
    return out


def build_example(force=False):
    """examples/c_host: a plain C++ / CUDA-runtime host linked against libalmpc.so (no Python, no torch)."""
    src = os.path.join(ROOT, "examples", "c_host.cpp")
    out = os.path.join(ROOT, "examples", "c_host")
    if not os.path.exists(src):
        return None
    if not force and os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(OUT)):
        return out
    cmd = [nvcc_path(), "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"), src, "-o", out,
           "-L", HERE, "-lalmpc", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../deq_mpc_corl_b200"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("example build failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    build(force=True, verbose="--verbose" in sys.argv)
    print("built", OUT)
    print("built", build_example(force=True))
