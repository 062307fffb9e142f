"""ctypes binding of libs2v_b200.so (the C ABI declared in include/s2v_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libs2v_b200.so")

EPI_BIAS, EPI_BIAS_GELU, EPI_GATE_RESIDUAL = 0, 1, 2


class LinearArgs(C.Structure):
    """s2v_linear_args (include/s2v_b200.h)."""

    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("bias", C.c_void_p),
        ("lora_t", C.c_void_p), ("ldt", C.c_int64),
        ("lora_b", C.c_void_p), ("ldb", C.c_int64),
        ("lora_r", C.c_int32), ("lora_group_n", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("epilogue", C.c_int32),
        ("alpha", C.c_float),
        ("mod", C.c_void_p),
        ("mod_stride", C.c_int32), ("gate_off_text", C.c_int32), ("gate_off_other", C.c_int32),
        ("rows_per_batch", C.c_int32), ("text_len", C.c_int32),
    ]


class SmallLinearDesc(C.Structure):
    """include/s2v_b200.h: s2v_small_linear_desc (one problem of s2v_small_linear_batch; the array lives in DEVICE memory)."""
    _fields_ = [("w", C.c_void_p), ("ldw", C.c_int64), ("bias", C.c_void_p), ("out", C.c_void_p), ("ldo", C.c_int64),
                ("x", C.c_void_p), ("ldx", C.c_int64), ("N", C.c_int32), ("K", C.c_int32), ("act_in", C.c_int32),
                ("round_bf16", C.c_int32), ("alpha", C.c_float), ("beta", C.c_float)]


class QkNormArgs(C.Structure):
    """s2v_qk_norm_args (include/s2v_b200.h)."""

    _fields_ = [
        ("nq_w", C.c_void_p), ("nq_b", C.c_void_p), ("nk_w", C.c_void_p), ("nk_b", C.c_void_p),
        ("cos", C.c_void_p), ("sin", C.c_void_p),
        ("S", C.c_int32), ("H", C.c_int32), ("text_len", C.c_int32),
        ("eps", C.c_float),
    ]


class ConvArgs(C.Structure):
    """s2v_conv_args (include/s2v_b200.h)."""

    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("bias", C.c_void_p),
        ("res", C.c_void_p), ("ldres", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("T", C.c_int32), ("t_pad", C.c_int32), ("Hp", C.c_int32), ("Wp", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("taps", C.c_int32),
    ]


_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

# name -> argtypes; every function returns int except s2v_last_error
SIGNATURES = {
    "s2v_abi_version": [],
    "s2v_device_check": [C.c_int],
    "s2v_linear": [C.POINTER(LinearArgs), _vp],
    "s2v_qkv_lora": [C.POINTER(LinearArgs), _vp],
    "s2v_qkv_lora_norm_rope": [C.POINTER(LinearArgs), C.POINTER(QkNormArgs), _vp],
    "s2v_outproj_lora_gate_residual": [C.POINTER(LinearArgs), _vp],
    "s2v_ffn_up_gelu_lora": [C.POINTER(LinearArgs), _vp],
    "s2v_ffn_down_lora_gate_residual": [C.POINTER(LinearArgs), _vp],
    "s2v_attn_fwd": [_vp, _vp, _i32, _i32, _i32, _f32, _vp],
    "s2v_adaln_modulate": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "s2v_final_norm": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "s2v_qk_norm_rope": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp],
    "s2v_small_linear": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _vp],
    "s2v_small_linear_batch": [_vp, _i32, _i32, _vp, _i64, _i32, _vp],
    "s2v_timestep_sinusoid": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "s2v_patchify": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "s2v_unpatchify": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "s2v_add_rows": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "s2v_cfg_ddim_step": [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _vp],
    "s2v_ddim_step": [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp],
    "s2v_dpm_step": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _vp],
    "s2v_conv_gemm": [C.POINTER(ConvArgs), _vp],
    "s2v_conv_gemm_stats": [C.POINTER(ConvArgs), _vp, C.POINTER(C.c_int32), _vp],
    "s2v_vae_groupnorm_finalize": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "s2v_vae_latent_rows": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "s2v_vae_latent_im2col": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "s2v_vae_groupnorm_stats": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "s2v_vae_spatialnorm_silu": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(C.c_int32), _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "s2v_vae_upsample_nearest": [_vp, _vp, C.POINTER(C.c_int32), _i32, _i32, _i32, _i32, _vp],
    "s2v_vae_volume_to_video": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "s2v_vae_groupnorm_silu": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "s2v_vae_subsample2": [_vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "s2v_video_to_uint8": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "s2v_gather_rows": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "s2v_rmsnorm": [_vp, _vp, _vp, _i32, _i32, _f32, _vp],
    "s2v_gated_gelu": [_vp, _vp, _i64, _i32, _vp],
    "s2v_t5_attention": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "s2v_vae_blend": [_vp, _vp, _i64, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _vp],
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built — there is no Python/CPU substitute."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(s2v_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch
            fn.argtypes = argtypes
            fn.restype = C.c_int
        lib.s2v_last_error.argtypes = []
        lib.s2v_last_error.restype = C.c_char_p
        lib.s2v_workspace_bytes.argtypes = [_i32] * 6
        lib.s2v_workspace_bytes.restype = C.c_int64
        if lib.s2v_abi_version() != 1:
            raise RuntimeError("libs2v_b200.so ABI version mismatch")
        _lib = lib
    return _lib


launch_count = 0  # kernels launched through the C ABI by this process (every compute entry point launches exactly one)


def check(rc: int, what: str):
    global launch_count
    launch_count += 1
    if rc != 0:
        msg = load().s2v_last_error().decode("utf-8", "replace")
        kind = "invalid argument / unsupported" if rc < 0 else "CUDA error"
        raise RuntimeError(f"{what} failed ({kind}, code {rc}): {msg}")
