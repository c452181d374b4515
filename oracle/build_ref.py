"""Compile the reference's OWN native sources, where they lie under /root/reference, into oracle/_ref/
(TEST INFRASTRUCTURE; outputs are git-ignored but travel to the GPU box).

  wsovod_ref_C.so   the reference extension exactly as its setup.py builds it (wsovod/layers/vision.cpp,
                    ROILoopPool/*.cpp|.cu, csc/*.cu; flags of setup.py:70-81) for sm_100a: the GPU-only
                    oracle of ROILoopPool forward/backward (its dispatcher rejects CPU tensors,
                    ROILoopPool.h:62).
  wsovod_ref_cpu.so a 12-line pybind shim (written here, not reference code) that exposes the
                    reference's ROILoopPool_forward_cpu / _backward_cpu (ROILoopPool_cpu.cpp:125-232),
                    which the reference compiles but never dispatches to: the CPU oracle of ROIPool.
  py/               the reference's hot-path PYTHON files, byte for byte (wsovod/modeling/roi_heads/roi_heads.py,
                    fast_rcnn_open_vocabulary.py, class_heads/open_vocabulary_classifier.py, poolers.py,
                    layers/roi_loop_pool.py, layers/csc.py): /root/reference does not exist on the GPU box, so the
                    drop-in test (tests/test_gpu_dropin.py) and `bench.py --impl reference` import the
                    reference from here through oracle/d2_shim.py.
No reference source is copied into the repository's history (oracle/_ref/ is git-ignored); the compiler reads the
native files in place.
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

SHIM = r'''
#include <torch/extension.h>
#include "ROILoopPool/ROILoopPool.h"
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("roi_pool_forward_cpu", &wsovod::ROILoopPool_forward_cpu);
  m.def("roi_pool_backward_cpu", &wsovod::ROILoopPool_backward_cpu);
}
'''


PY_FILES = [
    "wsovod/modeling/roi_heads/roi_heads.py",
    "wsovod/modeling/roi_heads/fast_rcnn_open_vocabulary.py",
    "wsovod/modeling/class_heads/open_vocabulary_classifier.py",
    "wsovod/modeling/poolers.py",
    "wsovod/layers/roi_loop_pool.py",
    "wsovod/layers/csc.py",
]


def ship_python(reference):
    for rel in PY_FILES:
        dst = os.path.join(OUT, "py", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference, rel), dst)
    print("shipped", len(PY_FILES), "reference python files to", os.path.join(OUT, "py"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--cpu-only", action="store_true")
    a = ap.parse_args()
    layers = os.path.join(a.reference, "wsovod", "layers")
    if not os.path.isdir(layers):
        print("reference not present; nothing to build")
        return 0
    os.makedirs(OUT, exist_ok=True)
    ship_python(a.reference)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    build = os.path.join("/tmp", "wsovod_ref_build")
    os.makedirs(build, exist_ok=True)
    shim = os.path.join(build, "cpu_shim.cpp")
    with open(shim, "w") as f:
        f.write(SHIM)
    os.makedirs(os.path.join(build, "cpu"), exist_ok=True)
    load(name="wsovod_ref_cpu", sources=[shim, os.path.join(layers, "ROILoopPool", "ROILoopPool_cpu.cpp")],
         extra_include_paths=[layers], extra_cflags=["-O2"], build_directory=os.path.join(build, "cpu"),
         is_python_module=False, verbose=False)
    shutil.copy(os.path.join(build, "cpu", "wsovod_ref_cpu.so"), os.path.join(OUT, "wsovod_ref_cpu.so"))
    print("built", os.path.join(OUT, "wsovod_ref_cpu.so"))
    if not a.cpu_only:
        os.makedirs(os.path.join(build, "cuda"), exist_ok=True)
        srcs = [os.path.join(layers, "vision.cpp"),
                os.path.join(layers, "ROILoopPool", "ROILoopPool_cpu.cpp"),
                os.path.join(layers, "ROILoopPool", "ROILoopPool_cuda.cu"),
                os.path.join(layers, "csc", "csc_cuda.cu")]
        load(name="wsovod_ref_C", sources=srcs, extra_include_paths=[layers], with_cuda=True,
             extra_cflags=["-DWITH_CUDA"],
             extra_cuda_cflags=["-O3", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__", "-DWITH_CUDA"],
             build_directory=os.path.join(build, "cuda"), is_python_module=False, verbose=False)
        shutil.copy(os.path.join(build, "cuda", "wsovod_ref_C.so"), os.path.join(OUT, "wsovod_ref_C.so"))
        print("built", os.path.join(OUT, "wsovod_ref_C.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
