"""ctypes binding of libfdn_b200.so (the C ABI declared in include/fdn_b200.h).

There is exactly one implementation behind these symbols: the sm_100a CUDA library built in-tree by
``fdn_tip2025_b200.build`` / ``__graft_entry__.build()``.  If it is missing the import of any op fails
loudly; there is no CPU or PyTorch fallback.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfdn_b200.so")

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float

# name -> argument type codes (p pointer, i int, l long long, f float, s cudaStream_t)
SIGNATURES = {
    "fdn_abi_version": "",
    "fdn_is_device_build": "",
    "fdn_fft_prepare": "ii",
    "fdn_fft_rows_r2c": "ppiiis",
    "fdn_fft_rows_c2r": "ppiiifpfpis",
    "fdn_fft_cols": "plipliiiiiiipppps",
    "fdn_spec_mlp": "plliips",
    "fdn_fdffn_patch": "ppppiiiis",
    "fdn_fdffn_patch_dw": "pppppiiiis",
    "fdn_fdffn_spatial": "pppppiiiis",
    "fdn_fdsa_patch": "pppiiiis",
    "fdn_fdsa_patch_dw": "pppppiiiis",
    "fdn_pw_conv": "piipiipiippppipppfpplliiiiis",
    "fdn_has_tcgen05": "",
    "fdn_pw_mma_set_debug": "p",
    "fdn_pw_mma_supported": "iiii",
    "fdn_pw_mma": "pipipiiiippplpppppfpiiis",
    "fdn_group_stats": "ppiiiis",
    "fdn_chan_ln": "ppppplpliiiis",
    "fdn_avgpool2": "ppiiis",
    "fdn_up2_bilinear": "ppiiis",
    "fdn_pixel_unshuffle": "ppiiiiis",
    "fdn_gamma_curve": "pppfls",
    "fdn_fill_border": "ppiiiis",
    "fdn_conv2d": "ppppipiiiiiiiiiis",
    "fdn_conv3x3_mma_cn": "i",
    "fdn_conv3x3_mma": "pppppiiiiis",
    "fdn_film_maps": "pppppiiiis",
    "fdn_convt4s2": "ppppiiiiiis",
    "fdn_dwconv3": "pppiiiiis",
    "fdn_avgpool3s2": "ppiiis",
    "fdn_plane_mean": "ppiis",
    "fdn_se_fc": "ppppppiiis",
    "fdn_se_apply": "ppppiis",
    "fdn_lpnet_head": "pppppppiis",
    "fdn_gray_mean": "ppiis",
    "fdn_pre_u8hwc_to_f32chw": "ppiiiiis",
    "fdn_post_f32chw_to_u8hwc": "ppiiiiis",
    "fdn_psnr": "ppppiiiiiis",
    "fdn_ssim": "ppppiiiiiis",
    "fdn_diff": "pppls",
    "fdn_reduce_f64": "ppplis",
    "fdn_down8_bilinear": "ppiiis",
}
_CODE = {"p": _P, "i": _I, "l": _L, "f": _F, "s": _P}

_handle = None


def _bind(handle):
    for name, codes in SIGNATURES.items():
        fn = getattr(handle, name)          # AttributeError if the library does not export the symbol
        fn.argtypes = [_CODE[c] for c in codes]
        fn.restype = _I
    handle.fdn_last_error_string.argtypes = []
    handle.fdn_last_error_string.restype = ctypes.c_char_p
    return handle


def load():
    """Return the bound library handle, loading libfdn_b200.so on first use."""
    global _handle
    if _handle is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libfdn_b200.so is not built (%s). Run `python -m fdn_tip2025_b200.build` "
                "(needs nvcc with sm_100a support); there is no CPU fallback." % LIB_PATH)
        h = _bind(ctypes.CDLL(LIB_PATH))
        if h.fdn_is_device_build() != 1:
            raise RuntimeError("libfdn_b200.so is not a device build")
        _handle = h
    return _handle


# launch accounting (bench.py reads these; one C-ABI call == one kernel launch, fdn_fft_prepare launches nothing)
launch_count = 0
profile_hook = None      # optional callable(name, fn) -> rc used by bench.py to time single launches with CUDA events


def call(name, *args):
    """Invoke a C-ABI entry point; non-zero return codes become RuntimeError(fdn_last_error_string())."""
    global launch_count
    h = load()
    fn = getattr(h, name)
    if profile_hook is not None:
        rc = profile_hook(name, lambda: fn(*args), args)
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, h.fdn_last_error_string().decode()))
    if name != "fdn_fft_prepare":
        launch_count += 1
