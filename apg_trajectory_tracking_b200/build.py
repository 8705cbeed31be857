"""Build libapg_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m apg_trajectory_tracking_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libapg_b200.so")
SOURCES = ["capi.cu", "capi_prep.cu", "prep_kernels.cu", "eval_kernels.cu", "learnt_kernels.cu", "misc_kernels.cu", "p2p_kernels.cu", "hutter_kernels.cu", "tq_kernels.cu", "tq_dw_kernels.cu", "simple_kernels.cu", "rec_kernels.cu", "lstm_kernels.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-Xcudafe", "--diag_suppress=177"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "apg_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, profile=False, variant=None, defines=()):
    """profile=True builds libapg_b200_prof.so with per-phase cycle counters (tools/tq_profile.py);
    variant="name" with defines=("-DX", ...) builds libapg_b200_<name>.so for A/B timing experiments (load it with
    APG_B200_LIB=<path>).  The translation units are compiled in parallel (one nvcc per .cu) and linked into one
    shared library."""
    from concurrent.futures import ThreadPoolExecutor
    if profile:
        variant, defines = "prof", ("-DAPG_PROFILE",) + tuple(defines)
    out = LIB.replace(".so", f"_{variant}.so") if variant else LIB
    if not force and not variant and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build", variant or "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--shared"] + (["-Xptxas", "-v"] if verbose else []) + list(defines)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        res = subprocess.run([_nvcc()] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True,
                             text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = [r for r in results if r[2].returncode != 0]
    for src, _, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
    if failed:
        raise RuntimeError("nvcc failed on " + ", ".join(r[0] for r in failed))
    link = subprocess.run([_nvcc(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a"] +
                          [r[1] for r in results] + ["-o", out], capture_output=True, text=True)
    if verbose or link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
    if link.returncode != 0:
        raise RuntimeError("nvcc failed linking libapg_b200.so")
    return out


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, profile="--profile" in sys.argv,
                variant=_variant, defines=tuple(a for a in sys.argv if a.startswith("-D"))))
