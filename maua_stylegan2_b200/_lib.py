"""ctypes binding of libmaua_b200.so (include/maua_b200.h).  Fails loudly when the library is missing:
there is NO CPU / eager fallback for CUDA tensors anywhere in this package."""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAUA_LIB_PATH") or os.path.join(_PKG, "libmaua_b200.so")   # MAUA_LIB_PATH: A/B builds (tools/)


class MauaError(RuntimeError):
    pass


class StyleJob(C.Structure):
    """MauaStyleJob (include/maua_b200.h)."""
    _fields_ = [("mod_w", C.c_void_p), ("mod_b", C.c_void_p), ("wsq", C.c_void_p), ("s_out", C.c_void_p),
                ("d_out", C.c_void_p), ("s_norm_out", C.c_void_p), ("cin", C.c_int32), ("cout", C.c_int32), ("latent_index", C.c_int32),
                ("reserved", C.c_int32)]


class ConvEpilogue(C.Structure):
    """MauaConvEpilogue (include/maua_b200.h)."""
    _fields_ = [("d", C.c_void_p), ("noise", C.c_void_p), ("noise_weight", C.c_void_p), ("bias", C.c_void_p),
                ("s_next", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_f32_nchw", C.c_void_p),
                ("out_raw_nhwc", C.c_void_p), ("noise_bstride", C.c_longlong), ("slope", C.c_float),
                ("act_scale", C.c_float), ("activate", C.c_int32), ("out_fmt", C.c_int32), ("rgb_w", C.c_void_p),
                ("rgb_out", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong)]


class SynthLayer(C.Structure):
    """MauaSynthLayer (include/maua_b200.h)."""
    _fields_ = [(n, C.c_void_p) for n in ("conv_weight", "mod_weight", "mod_bias", "noise_weight", "act_bias",
                                          "noise_buffer", "blur_kernel", "rgb_weight", "rgb_mod_weight", "rgb_mod_bias",
                                          "rgb_bias", "rgb_up_kernel")] + \
               [(n, C.c_int32) for n in ("cin", "cout", "up", "latent_index", "rgb_latent_index", "reserved")]


class SynthDesc(C.Structure):
    """MauaSynthDesc (include/maua_b200.h)."""
    _fields_ = [("layers", C.POINTER(SynthLayer)), ("const_input", C.c_void_p)] + \
               [(n, C.c_int32) for n in ("n_layers", "in_h", "in_w", "style_dim", "n_latent", "precision", "f16_min_res",
                                         "min_rgb_size")]


_p, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes; every function returns int (0 ok / negative error) unless listed in _SPECIAL
SIGNATURES = {
    "maua_upfirdn2d_f32": [_p, _p, _p] + [_i] * 14 + [_p],
    "maua_fused_bias_act_f32": [_p, _p, _p, _p, _ll, _i, _i, _i, _i, _f, _f, _p],
    "maua_linear_f32": [_p, _p, _p, _p, _i, _i, _i, _f, _f, _i, _i, _p],
    "maua_style_prologue_f32": [_p, _i, _p, _p, _p, _f, _p, _i, _i, _i, _p],
    "maua_weight_sq_f32": [_p, _p, _i, _i, _i, _f, _p],
    "maua_modconv_simt_f32": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p],
    "maua_noise_bias_act_f32": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _ll, _f, _f, _p],
    "maua_torgb_f32": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p],
    "maua_rgb_to_u8_nhwc": [_p, _p, _i, _i, _i, _p],
    "maua_rgb_weights_f32": [_p, _p, _p, _i, _i, _f, _p],
    "maua_rgb_finish_f32": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "maua_rgb_finish_u8": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "maua_pack_weight_bf16x2": [_p, _p, _p, _i, _i, _i, _f, _p],
    "maua_pack_weight_f16x2": [_p, _p, _p, _i, _i, _i, _f, _p],
    "maua_modulate_split_nhwc": [_p, _ll, _p, _p, _p, _i, _i, _i, _i, _p],
    "maua_modulate_f16_nhwc": [_p, _ll, _p, _p, _i, _i, _i, _i, _p],
    "maua_modconv_tc": [_p, _p, _p, _p, C.POINTER(ConvEpilogue), _i, _i, _i, _i, _i, _i, _i, _p],
    "maua_blur_act_nhwc": [_p, _p, C.POINTER(ConvEpilogue), _i, _i, _i, _i, _p],
    "maua_audio_stft_f32": [_p, _ll, _p, _p, _i, _i, _i, _p],
    "maua_audio_istft_f32": [_p, _p, _ll, _p, _i, _i, _i, _p],
    "maua_audio_hpss_f32": [_p, _p, _p, _i, _i, _f, _f, _i, _p],
    "maua_audio_filterbank_f32": [_p, _p, _p, _i, _i, _i, _p],
    "maua_audio_onset_env_f32": [_p, _p, _p, _i, _i, _i, _f, _f, _p],
    "maua_audio_stft_mm_f32": [_p, _ll, _p, _p, _i, _i, _i, _p],
    "maua_audio_onsets_mm_f32": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "maua_audio_rms_f32": [_p, _p, _i, _i, _i, _p],
    "maua_audio_cens_f32": [_p, _p, _p, _i, _i, _i, _p],
    "maua_audio_nn_filter_f32": [_p, _p, _p, _i, _i, _i, _i, _p],
    "maua_resample_f32": [_p, _p, _p, _i, _i, _i, _p],
    "maua_clip_to_range_f32": [_p, _i, _p, _i, _p],
    "maua_gaussian_filter_f32": [_p, _p, _i, _ll, _f, _f, _i, _f, _p],
    "maua_percentile_clip_f32": [_p, _p, _i, _f, _f, _p],
    "maua_sosfilt_f32": [_p, _p, _ll, _p, _i, _p],
    "maua_chroma_weight_latents_f32": [_p, _p, _p, _i, _i, _ll, _p],
    "maua_envelope_blend_f32": [_p, _p, _p, _i, _ll, _p],
    "maua_fit_frames_u8": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _i, _p],
    "maua_bend_warp_f32": [_p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "maua_perlin_noise": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "maua_synth_create": [C.POINTER(SynthDesc), C.POINTER(C.c_void_p)],
    "maua_synth_prepare": [_p, _p, C.c_size_t, _p],
    "maua_synth_bind": [_p, _p, C.c_size_t, _i, _p],
    "maua_synth_forward": [_p, _p, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), _p, _p, _f, _i, _p, _p, _p],
}
_SPECIAL = {
    "maua_abi_version": (C.c_int, []),
    "maua_last_error": (C.c_char_p, []),
    "maua_launch_count": (C.c_longlong, []),
    "maua_modconv_tc_last_config": (C.c_char_p, []),
    "maua_synth_destroy": (None, [C.c_void_p]),
    "maua_synth_plan_bytes": (C.c_size_t, [C.c_void_p]),
    "maua_synth_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "maua_synth_truncated_latents": (C.c_void_p, [C.c_void_p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MauaError(f"{LIB_PATH} is missing: run `python -m maua_stylegan2_b200.build` "
                            "(or __graft_entry__.build()); there is no fallback path")
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        for name, (res, argtypes) in _SPECIAL.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = res
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        raise MauaError(f"{what} failed (rc={rc}): {lib().maua_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def launch_count():
    return int(lib().maua_launch_count())


def last_conv_config():
    """Variant / tile configuration of this thread's last maua_modconv_tc call (diagnostics, tests)."""
    return lib().maua_modconv_tc_last_config().decode()


# Optional per-launch timing (bench.py's roofline pass): PROFILE = {"names": set, "events": []}; TAG is attached to
# every recorded launch by the caller (synthesis.py sets it to the layer's algorithmic work).
PROFILE = None
TAG = None


def call(name, *args):
    if PROFILE is not None and name in PROFILE["names"]:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        check(getattr(lib(), name)(*args), name)
        end.record()
        PROFILE["events"].append((name, TAG, start, end))
        return
    check(getattr(lib(), name)(*args), name)
