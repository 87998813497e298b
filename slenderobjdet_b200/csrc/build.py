"""Build libslender_b200.so in-tree with nvcc for sm_100a (no torch headers, no libcuda link).

    python -m slenderobjdet_b200.csrc.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SO = os.path.join(PKG, "libslender_b200.so")
SOURCES = ["api.cu", "dcn_simt.cu", "dcn_tc.cu", "dcn_tc_bwd.cu", "dcn_tc_dx.cu", "dcn_tf32.cu", "gn.cu", "assign.cu", "losses.cu", "postproc.cu"]
HEADERS = ["common.cuh", "tc_common.cuh", "dcn_tc_shared.cuh", os.path.join("..", "..", "include", "slender_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _host_compiler_flags():
    # the image exports CC/CXX=/opt/gcc/bin/*, fine for nvcc's host pass; let nvcc pick its default
    return []


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    deps = srcs + [os.path.join(HERE, h) for h in HEADERS if os.path.exists(os.path.join(HERE, h))]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    newest_dep_h = max(os.path.getmtime(d) for d in deps if not d.endswith(".cu"))
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if (not force and os.path.exists(o) and os.path.getmtime(o) >= os.path.getmtime(s)
                and os.path.getmtime(o) >= newest_dep_h):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + _host_compiler_flags() + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((s, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("FAILED: %s\n%s\n" % (" ".join(cmd), out))
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed (see stderr above)")
    if procs or force or not os.path.exists(SO):
        cmd = [_nvcc(), "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-fPIC"]
        subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
