"""The gather kernels of csrc/gradient.cu executed on the HOST: their source text is extracted verbatim from the .cu / .cuh
files and compiled with g++ against tests/emu/cuda_emu.h (threads = OS threads, warp shuffles and barriers = rendezvous,
__shared__ = per-CTA static).  The kernel variants written without GPU access (software-pipelined, bundle-staged,
batched-gather; options "spmm_prefetch" 1/2/3, "spmm_pad") must reproduce the default kernels' rows bit for bit and their
fused sums to rounding; the default kernels are checked against a plain CSR product.  This does not replace the GPU parity
tests -- it makes sure the first GPU minutes are not spent on indexing mistakes."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sdplrplus.jl_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")


def _between(text, start, end):
    i = text.index(start)
    j = text.index(end, i)
    return text[i:j]


def extract(build_dir):
    common = open(os.path.join(CSRC, "common.cuh")).read()
    grad = open(os.path.join(CSRC, "gradient.cu")).read()
    # device helpers: warp_sum .. ldg2 (everything between the "device helpers" banner and the launcher prototypes)
    helpers = _between(common, "__device__ __forceinline__ double warp_sum", "// host-visible kernels' launcher prototypes")
    # Acc, RowArgs, epilogues and every row kernel: from the Acc declaration to the first non-gather kernel after them
    kernels = _between(grad, "template <int VEC>\nstruct Acc;", "__global__ void k_lr_apply(")
    assert "<<<" not in kernels and "<<<" not in helpers, "host launches inside the extracted device code"
    for name in ("k_rows_group", "k_rows_warp", "k_rows_combine", "k_rows_group_pf", "k_rows_warp_pf", "k_rows_bundle", "k_rows_group_b",
                 "owned_q_range", "row_epilogue", "finish_sums"):
        assert name in kernels, name
    open(os.path.join(build_dir, "extracted_common.inc"), "w").write(helpers)
    open(os.path.join(build_dir, "extracted_gradient.inc"), "w").write(kernels)


def extract_lanczos(build_dir):
    common = open(os.path.join(CSRC, "common.cuh")).read()
    lz = open(os.path.join(CSRC, "lanczos.cu")).read()
    helpers = _between(common, "__device__ __forceinline__ double warp_sum", "// host-visible kernels' launcher prototypes")
    kernels = _between(lz, "template <int LANES, bool WITH_ALPHA>", "}  // namespace")   # every kernel of the q-step path
    assert "<<<" not in kernels
    for name in ("k_lz_spmv", "k_lz_spmv_bundle", "k_lz_class_ranges", "k_lz_update_part", "k_lz_beta_finish", "k_lz_update"):
        assert name in kernels, name
    open(os.path.join(build_dir, "extracted_common.inc"), "w").write(helpers)
    open(os.path.join(build_dir, "extracted_lanczos.inc"), "w").write(kernels)


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_lanczos_kernels_under_host_emulation(tmp_path):
    """bundle SpMV (option "lanczos_bundle") and the kernels of the row-partitioned Lanczos (option "lanczos_dist")."""
    extract_lanczos(str(tmp_path))
    exe = os.path.join(str(tmp_path), "lanczos_emu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-I", EMU, "-I", str(tmp_path), "-o", exe,
                           os.path.join(EMU, "lanczos_emu_main.cpp")])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "the Lanczos kernels agree" in res.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_gather_kernel_variants_under_host_emulation(tmp_path):
    extract(str(tmp_path))
    exe = os.path.join(str(tmp_path), "gather_emu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-I", EMU, "-I", str(tmp_path), "-o", exe,
                           os.path.join(EMU, "gather_emu_main.cpp")])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "all kernel variants agree" in res.stdout
