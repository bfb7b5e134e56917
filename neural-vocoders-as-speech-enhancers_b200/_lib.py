"""ctypes binding of include/nvse_b200.h (csrc/libnvse_b200.so).

There is no CPU fallback: if the shared library is missing or no CUDA device is
present, every compute entry point raises.  Build the library in-tree with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C csrc``)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NVSE_LIB: an alternative build of the same library (same-box A/B timing of kernel variants, tools/ab_build.sh)
LIB_PATH = os.environ.get("NVSE_LIB") or os.path.join(_HERE, "csrc", "libnvse_b200.so")

NVSE_OK = 0
PRECISION_F32, PRECISION_BF16 = 0, 1
GEN_HIFIGAN, GEN_ISTFTNET = 0, 1
MAX_UPS = MAX_KERNELS = MAX_DILATIONS = 8
ABI_VERSION = 1


class NvseError(RuntimeError):
    """A library call returned non-zero; the message is nvse_last_error()."""


class GeneratorConfig(C.Structure):
    """struct nvse_generator_config (include/nvse_b200.h)."""

    _fields_ = [
        ("kind", C.c_int32), ("in_channels", C.c_int32), ("initial_channel", C.c_int32),
        ("num_upsamples", C.c_int32), ("upsample_rates", C.c_int32 * MAX_UPS),
        ("upsample_kernel_sizes", C.c_int32 * MAX_UPS), ("resblock_type", C.c_int32),
        ("num_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * MAX_KERNELS),
        ("num_dilations", C.c_int32 * MAX_KERNELS),
        ("resblock_dilations", (C.c_int32 * MAX_DILATIONS) * MAX_KERNELS),
        ("istft_n_fft", C.c_int32), ("istft_hop", C.c_int32),
    ]


_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/nvse_b200.h declares
PROTOTYPES = {
    "nvse_abi_version": (_i, []),
    "nvse_last_error": (C.c_char_p, []),
    "nvse_launch_count": (C.c_uint64, []),
    "nvse_profile_begin": (_i, []),
    "nvse_profile_end": (_i, [C.c_char_p, _sz]),
    "nvse_frontend_create": (_i, [_i, _i, _i, _vp, _vp, C.POINTER(_vp)]),
    "nvse_frontend_destroy": (_i, [_vp]),
    "nvse_frontend_num_frames": (_i64, [_vp, _i64]),
    "nvse_frontend_mel_f32": (_i, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "nvse_frontend_mel_ragged_f32": (_i, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "nvse_frontend_stft_f32": (_i, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "nvse_inverse_mel_f32": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i64, _vp]),
    "nvse_frontend_backward_scratch_bytes": (_sz, [_vp, _i64, _i64]),
    "nvse_frontend_istft_scratch_bytes": (_sz, [_vp, _i64, _i64]),
    "nvse_frontend_istft_f32": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "nvse_frontend_istft_c64": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "nvse_frontend_mel_backward_f32": (_i, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "nvse_generator_create": (_i, [C.POINTER(GeneratorConfig), C.POINTER(_vp)]),
    "nvse_generator_destroy": (_i, [_vp]),
    "nvse_generator_set_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _vp]),
    "nvse_generator_finalize": (_i, [_vp, _vp]),
    "nvse_generator_num_layers": (_i, [_vp]),
    "nvse_generator_layer_name": (_i, [_vp, _i, C.c_char_p, _sz]),
    "nvse_generator_load_weights": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i, _i, _vp]),
    "nvse_generator_out_samples": (_i64, [_vp, _i64]),
    "nvse_generator_workspace_bytes": (_sz, [_vp, _i64, _i64, _i]),
    "nvse_generator_forward": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _i, _vp]),
    "nvse_generator_forward_pcm16": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _i, _vp]),
    "nvse_generator_forward_ragged": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _sz, _i, _vp]),
    "nvse_weight_norm_fold_f32": (_i, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "nvse_pcm16_from_f32": (_i, [_vp, _vp, _i64, _vp]),
    "nvse_transpose_bct_to_btc_f32": (_i, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "nvse_transpose_btc_to_bct_f32": (_i, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "nvse_conv1d_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "nvse_conv1d_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "nvse_resblock1_bf16": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                 C.POINTER(_i), _i, _vp, _i64, _i64, _i, _i, _f, _i, _vp]),
    "nvse_conv_transpose1d_f32": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _i, _f, _vp]),
    "nvse_conv_transpose1d_bf16": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _i, _f, _vp]),
    "nvse_istft_head_f32": (_i, [_vp, _vp, _i64, _i64, _i, _i, _vp]),
    "nvse_generator_tape_bytes": (_sz, [_vp, _i64, _i64]),
    "nvse_generator_backward_workspace_bytes": (_sz, [_vp, _i64, _i64]),
    "nvse_generator_grad_elems": (_i64, [_vp]),
    "nvse_generator_grad_offset": (_i, [_vp, C.c_char_p, C.POINTER(_i64), C.POINTER(_i64)]),
    "nvse_generator_forward_train": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _i, _vp]),
    "nvse_generator_backward": (_i, [_vp, _i64, _i64, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _i, _vp]),
    "nvse_generator_total_rows": (_i64, [_vp]),
    "nvse_generator_weight_norm_backward": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _i, _vp]),
    "nvse_conv1d_backward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _f, _i, _vp]),
    "nvse_conv_transpose1d_backward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _i, _f, _vp]),
    "nvse_weight_norm_backward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "nvse_istft_head_backward_f32": (_i, [_vp, _vp, _vp, _i64, _i64, _i, _i, _vp]),
    "nvse_vocoder_mel_pitch": (_i, [_vp]),
    "nvse_vocoder_workspace_bytes": (_sz, [_vp, _vp, _i64, _i64]),
    "nvse_vocoder_forward": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nvse_disc_conv_out_len": (_i64, [_i64, _i, _i, _i]),
    "nvse_disc_conv_forward_f32": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i64, _i, _i, _i, _i, _i, _f, _vp]),
    "nvse_disc_conv_backward_scratch_bytes": (_sz, [_i64, _i, _i, _i64, _i, _i, _i, _i, _i]),
    "nvse_disc_conv_backward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i64, _i, _i, _i, _i, _i, _f, _vp, _sz, _vp]),
    "nvse_avgpool1d_f32": (_i, [_vp, _vp, _i64, _i64, _i, _i, _i, _vp]),
    "nvse_avgpool1d_backward_f32": (_i, [_vp, _vp, _i64, _i64, _i, _i, _i, _vp]),
    "nvse_tc_abort_status": (_i, [_i, C.POINTER(_i)]),
    "nvse_debug_rb_trace": (_i, [C.POINTER(C.c_longlong)]),
}

_lib = None


def load():
    """dlopen the library once and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NvseError(
            f"{LIB_PATH} is not built; run `make -C {os.path.dirname(LIB_PATH)}` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = restype, argtypes
    got = lib.nvse_abi_version()
    if got != ABI_VERSION:
        raise NvseError(f"libnvse_b200.so ABI version {got}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    if rc != NVSE_OK:
        msg = load().nvse_last_error()
        raise NvseError(f"nvse_b200 error {rc}: {msg.decode(errors='replace') if msg else '?'}")


def launch_count():
    return int(load().nvse_launch_count())


def profile_begin():
    check(load().nvse_profile_begin())


def profile_end():
    """Per-kernel totals since profile_begin(): list of dicts (kernel, launches, ms, flops, bytes)."""
    import json
    buf = C.create_string_buffer(1 << 16)
    check(load().nvse_profile_end(buf, len(buf)))
    return json.loads(buf.value.decode())


def tc_abort_status(reset=True):
    """True if a tensor-core kernel tripped its bounded-wait timeout (results are then invalid)."""
    flag = C.c_int(0)
    check(load().nvse_tc_abort_status(1 if reset else 0, C.byref(flag)))
    return bool(flag.value)


def ptr(t):
    """Device (or host) address of a torch tensor as a c_void_p, None-safe."""
    return None if t is None else C.c_void_p(t.data_ptr())
