"""Compile the UNMODIFIED reference kernels into oracle/_ref/ (test infrastructure).

The reference is a torch C++/CUDA extension; its own build system (setup.py) is not used.  This
recipe compiles the sources where they lie under /root/reference/src with nvcc for sm_100a through
torch.utils.cpp_extension.load (ninja), producing two Python extension modules:

  oracle/_ref/MDCONV_REF_M2D*.so  <- src/mdeformable_conv.cu   (its own PYBIND11_MODULE: the
                                    modulated 2D forward/backward, the only pair HEAD registers)
  oracle/_ref/MDCONV_REF_X6*.so   <- src/deformable_conv.cu, src/deformable_conv3d.cu,
                                    src/mdeformable_conv3d.cu, src/utils.cu + oracle/ref_shim.cpp
                                    (declarations + registration of the six unregistered entries)

Outputs only go to oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).  Needs
/root/reference, so it runs in the build container only; the GPU box uses the prebuilt files.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MDCONV_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-D__CUDA_NO_HALF_OPERATORS__",
              "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__", "-w"]


def built():
    if not os.path.isdir(OUT):
        return False
    names = os.listdir(OUT)
    return any(n.startswith("MDCONV_REF_M2D") and n.endswith(".so") for n in names) and \
        any(n.startswith("MDCONV_REF_X6") and n.endswith(".so") for n in names)


def build(verbose=False):
    if built():
        return True
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    for name, sources in (
            ("MDCONV_REF_M2D", [os.path.join(src, "mdeformable_conv.cu"), os.path.join(src, "utils.cu")]),
            ("MDCONV_REF_X6", [os.path.join(src, f) for f in
                               ("deformable_conv.cu", "deformable_conv3d.cu", "mdeformable_conv3d.cu", "utils.cu")]
             + [os.path.join(HERE, "ref_shim.cpp")])):
        bd = os.path.join(OUT, "build_" + name)
        os.makedirs(bd, exist_ok=True)
        try:
            load(name=name, sources=sources, extra_cuda_cflags=NVCC_FLAGS, extra_cflags=["-O2", "-w"],
                 build_directory=bd, verbose=verbose, is_python_module=False)
        except Exception as e:  # is_python_module=False still dlopens; without a GPU that is fine, but be safe
            so = os.path.join(bd, name + ".so")
            if not os.path.exists(so):
                raise
            print("note:", type(e).__name__, e, file=sys.stderr)
        so = os.path.join(bd, name + ".so")
        os.replace(so, os.path.join(OUT, name + ".so"))
        shutil.rmtree(bd, ignore_errors=True)  # objects are not needed on the GPU box
    return built()


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("reference build:", "ok" if ok else "unavailable (no /root/reference?)")
    sys.exit(0 if ok else 1)
