"""Kernel-source checks on the CUDA-semantics emulator (tests/emu): the same .cu files compiled with g++ and run on
CPU tensors, compared with torch fp64 / the fp64 oracle.  This is a debugging aid for the GPU-less build container and
covers host-side index logic; the parity claims are made by tests/test_gpu_parity.py on the B200."""
import pytest

import parity_cases as P


@pytest.fixture(scope="module")
def emu():
    import os
    from emu import harness
    harness.enable()
    old = os.environ.get("FDN_B200_GEMM")
    os.environ["FDN_B200_GEMM"] = "ffma"      # the emulation build has no tcgen05 kernel and the package never downgrades silently
    yield "cpu"
    if old is None:
        os.environ.pop("FDN_B200_GEMM", None)
    else:
        os.environ["FDN_B200_GEMM"] = old
    # restore the product loader so later tests see the real library
    from fdn_tip2025_b200 import _lib, ops
    import importlib
    _lib._handle = None
    importlib.reload(ops)


@pytest.mark.parametrize("h,w", [(8, 8), (16, 24), (30, 14), (13, 22), (34, 26), (46, 58), (23, 62)])
def test_emu_fft(emu, h, w):
    P.case_rfft2_irfft2(emu, h, w)
    P.case_irfft2_nonhermitian(emu, h, w)


@pytest.mark.parametrize("h,w", [(16, 24), (64, 64), (46, 58)])
def test_emu_fcaffn_fft_stage(emu, h, w):
    P.case_fcaffn_fft_stage(emu, h, w)
    P.case_fcaffn_fft_stage(emu, h, w, big_phase=True)


def test_emu_fft_fast_and_prime_first(emu):
    P.case_rfft2_irfft2(emu, 64, 64, planes=1)        # register-pipeline kernels (fft_fast.cuh)
    P.case_rfft2_irfft2(emu, 46, 94, planes=1)        # 46 = 2*23, 47: prime radices first, no input twiddles
    P.case_rfft2_irfft2(emu, 58, 1334, planes=1)      # 667 = 23*29: first-pass and general prime passes in one transform
    P.case_rfft2_irfft2(emu, 24, 64, planes=1)        # 24 rows, 16 rows per CTA: the non-persistent row kernels (no whole tiles)
    P.case_rfft2_irfft2(emu, 160, 64, planes=3)       # persistent TMA-staged rows looping over several tiles per CTA, 16-column tiles


def test_emu_pointwise_and_convs(emu):
    P.case_pw_conv(emu)
    P.case_conv2d(emu)
    P.case_convt_dw_misc(emu)


def test_emu_transformer_block(emu):
    P.case_tblock(emu, 32, 8, 16, True, True)


def test_emu_fuse(emu):
    P.case_fuse_resample(emu, 32, 8, 8)


@pytest.mark.parametrize("fused", ["1", "0"])
def test_emu_fdffn_fused_variant(emu, monkeypatch, fused):
    monkeypatch.setenv("FDN_B200_FDFFN_FUSED", fused)
    P.case_tblock(emu, 32, 8, 16, False, False, seed=21)
    P.case_tblock(emu, 8, 24, 40, False, False, seed=22, b=1)


def test_emu_image_pre_post(emu):
    P.case_imgio(emu)
    P.case_imgio(emu, h=33, w=64, b=1)
