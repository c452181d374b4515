"""Loader for oracle/_ref (the reference's own native code compiled by oracle/build_ref.py).
TEST INFRASTRUCTURE ONLY.  Returns None when the prebuilt file is absent."""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    path = os.path.join(_HERE, "_ref", name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu():
    """module with roi_pool_forward_cpu / roi_pool_backward_cpu (ROILoopPool_cpu.cpp:125-232)"""
    return _load("wsovod_ref_cpu")


def cuda():
    """the reference extension: roi_loop_pool_forward / roi_loop_pool_backward / csc_forward"""
    return _load("wsovod_ref_C")
