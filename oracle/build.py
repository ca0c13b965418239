"""Build recipe for the oracle's native parts (test infrastructure).

build_native(): gcc-compiles oracle/native/oracle_native.c -> oracle/native/liboracle_native.so
build_ref():    when /root/reference is present (development container only), compiles the
                reference's OWN first-party native sources from where they lie -- nothing is copied
                into this repository -- into oracle/_ref/ (git-ignored, travels to the GPU box):
                  * models/utils/src/Array_Index.cpp                     -> _ref/Array_Index*.so
                  * models/bbox_post_process/src/{iou3d_cpu,iou3d_nms,iou3d_nms_api}.cpp
                    + iou3d_nms_kernel.cu (nvcc, sm_100a)                -> _ref/iou3d_nms_cuda*.so
                The reference's own build system (setup.py / CMake) is not run.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
REF_OUT = os.path.join(HERE, "_ref")
NATIVE_SRC = os.path.join(HERE, "native", "oracle_native.c")
NATIVE_LIB = os.path.join(HERE, "native", "liboracle_native.so")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + r.stderr[-4000:])
        raise RuntimeError("command failed: %s" % cmd[0])


def build_native(force=False):
    if not force and os.path.exists(NATIVE_LIB) and os.path.getmtime(NATIVE_LIB) >= os.path.getmtime(NATIVE_SRC):
        return NATIVE_LIB
    # -fopenmp: the kernel-map neighbour table runs its offsets in parallel (as MinkowskiEngine's CPU backend does)
    _run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", NATIVE_SRC, "-o", NATIVE_LIB, "-lm"])
    return NATIVE_LIB


def ref_paths():
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(REF_OUT, "Array_Index" + ext), os.path.join(REF_OUT, "iou3d_nms_cuda" + ext)


def build_ref(force=False):
    """returns (array_index_so, iou3d_so) or None when the reference is not available."""
    ai, iou = ref_paths()
    if not os.path.isdir(REF):
        return (ai, iou) if os.path.exists(ai) and os.path.exists(iou) else None
    os.makedirs(REF_OUT, exist_ok=True)
    import pybind11
    pyinc = sysconfig.get_paths()["include"]
    if force or not os.path.exists(ai):
        _run(["g++", "-O2", "-fPIC", "-shared", "-fopenmp", "-std=c++17", "-I", pybind11.get_include(), "-I", pyinc,
              os.path.join(REF, "models/utils/src/Array_Index.cpp"), "-o", ai])
    if force or not os.path.exists(iou):
        import torch
        from torch.utils import cpp_extension as ce
        src = os.path.join(REF, "models/bbox_post_process/src")
        tmp = os.path.join(REF_OUT, "obj")
        os.makedirs(tmp, exist_ok=True)
        incs = []
        for i in ce.include_paths():
            incs += ["-I", i]
        incs += ["-I", pyinc, "-I", "/usr/local/cuda/include"]
        defs = ["-DTORCH_EXTENSION_NAME=iou3d_nms_cuda", "-DTORCH_API_INCLUDE_EXTENSION_H",
                "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
        objs = []
        for f in ("iou3d_cpu.cpp", "iou3d_nms.cpp", "iou3d_nms_api.cpp"):
            o = os.path.join(tmp, f + ".o")
            _run(["g++", "-O2", "-fPIC", "-std=c++17", "-w", *defs, *incs, "-c", os.path.join(src, f), "-o", o])
            objs.append(o)
        o = os.path.join(tmp, "iou3d_nms_kernel.cu.o")
        _run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-Xcompiler", "-fPIC",
              "-c", os.path.join(src, "iou3d_nms_kernel.cu"), "-o", o])
        objs.append(o)
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        _run(["g++", "-shared", *objs, "-o", iou, "-L", tlib, "-L", "/usr/local/cuda/lib64", "-lc10", "-ltorch_cpu",
              "-ltorch", "-ltorch_python", "-lcudart", "-Wl,-rpath," + tlib])
    return ai, iou


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))
