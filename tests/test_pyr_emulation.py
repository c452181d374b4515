"""Host emulation of the block-max pooling geometry (wsovod_b200/csrc/pool_pyr.cuh: per-proposal block
classes, bin descriptors, plane recipes) against a brute-force scan with the reference's bin edges
(ROILoopPool_cpu.cpp:29-79).  Runs on CPU: the header is shared with the CUDA kernels."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_blockmax_geometry_emulation(tmp_path):
    exe = str(tmp_path / "pyr_emul")
    src = os.path.join(ROOT, "tests", "host", "pyr_emul.cpp")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-w", "-o", exe, src])
    out = subprocess.run([exe, "200"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_separable_roi_align_tables_emulation(tmp_path):
    """wsovod_b200/csrc/align_sep.cuh (the tap tables of roi_align7_sep_kernel) against torchvision's per-sample loop:
    footprint == touched cells for the adaptive grid, values within 1e-5, lists inside their capacity"""
    exe = str(tmp_path / "align_sep_emul")
    src = os.path.join(ROOT, "tests", "host", "align_sep_emul.cpp")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-w", "-o", exe, src])
    out = subprocess.run([exe, "30"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
