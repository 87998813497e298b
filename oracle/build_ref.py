"""Build oracle/_ref/_ref_C*.so: the REFERENCE's own deformable-convolution CUDA code, compiled for sm_100a from the
sources where they lie under /root/reference -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

    python oracle/build_ref.py [--force]

What is compiled (nothing is copied into this repo; the reference's build system is not run):
    /root/reference/detectron2/detectron2/layers/csrc/deformable/deform_conv_cuda.cu          (host launchers :272-1129)
    /root/reference/detectron2/detectron2/layers/csrc/deformable/deform_conv_cuda_kernel.cu   (kernels :96-1066)
    oracle/ref_binding.cpp                                                                    (pybind names of vision.cpp:76-92)
Outputs go to oracle/_ref/ only (git-ignored, NOT gpurun-ignored: the .so travels to the GPU box, which has no
/root/reference).  Uses: the fp32 GPU-vs-GPU pin of DCN backward / DCNv2 / fractional offsets
(tests/test_gpu_ref_c.py) and the optional "reference CUDA kernels on the same B200" timing column of bench.py.
The reference has no CPU implementation of this path, so this extension cannot serve as a CPU oracle.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
CSRC = "/root/reference/detectron2/detectron2/layers/csrc"
NAME = "_ref_C"


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(force=False):
    """Returns the path of the extension, or None when /root/reference is absent (GPU box: prebuilt file only)."""
    so = so_path()
    srcs = [os.path.join(CSRC, "deformable", "deform_conv_cuda.cu"),
            os.path.join(CSRC, "deformable", "deform_conv_cuda_kernel.cu"),
            os.path.join(HERE, "ref_binding.cpp")]
    if not all(os.path.exists(s) for s in srcs[:2]):
        return so if os.path.exists(so) else None
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    from torch.utils import cpp_extension as ce
    import torch
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = ["-I" + p for p in ce.include_paths(device_type="cuda") + [sysconfig.get_paths()["include"], CSRC]]
    defs = ["-DWITH_CUDA", "-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(s).rsplit(".", 1)[0] + ".o")
        objs.append(o)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
               "--expt-relaxed-constexpr", "-w"] + (["-x", "cu"] if s.endswith(".cpp") else []) + defs + inc + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write("FAILED: %s\n%s\n" % (" ".join(cmd), out))
            raise RuntimeError("building the reference extension failed")
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = [nvcc, "-shared", "-o", so] + objs + ["-L" + tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                                                 "-ltorch_python", "-Xlinker", "-rpath", "-Xlinker", tlib]
    subprocess.run(link, check=True)
    for o in objs:
        os.remove(o)
    return so


def load():
    """Import the built extension (torch must be imported first); None when it was never built."""
    so = so_path()
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
