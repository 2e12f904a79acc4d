"""ctypes binding of libb200whisper.so (the C ABI in include/b200_whisper.h).

PyTorch owns every device buffer; this module only turns tensors into raw pointers and checks status
codes.  There is NO fallback: if the CUDA library cannot be loaded or a call fails, a RuntimeError is
raised (the product path never routes through a CPU implementation).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200W_LIB") or os.path.join(_HERE, "csrc", "libb200whisper.so")  # B200W_LIB: A/B builds

_lock = threading.Lock()
_lib = None

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_longlong
f32 = C.c_float
u64 = C.c_ulonglong
sz = C.c_size_t


class LogmelTables(C.Structure):
    _fields_ = [("hann", vp), ("tw400", vp)]


class FilterParams(C.Structure):
    _fields_ = [("n_vocab", i32), ("logits_ld", i32), ("sample_begin", i32), ("eot", i32), ("blank", i32),
                ("no_timestamps", i32), ("timestamp_begin", i32), ("no_speech", i32),
                ("max_initial_timestamp_index", i32), ("apply_timestamp_rules", i32), ("suppress_blank", i32),
                ("tokens_ld", i32), ("temperature", f32), ("seed", u64)]


class Dims(C.Structure):
    _fields_ = [(n, i32) for n in ("n_mels", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
                                   "n_vocab", "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer")]


class EncLayer(C.Structure):
    _fields_ = [(n, vp) for n in ("attn_ln_g", "attn_ln_b", "w_qkv", "b_qkv", "w_out", "b_out", "mlp_ln_g",
                                  "mlp_ln_b", "w_mlp1", "b_mlp1", "w_mlp2", "b_mlp2")]


class DecLayer(C.Structure):
    _fields_ = [(n, vp) for n in ("attn_ln_g", "attn_ln_b", "w_qkv", "b_qkv", "w_out", "b_out", "cross_ln_g",
                                  "cross_ln_b", "w_cq", "b_cq", "w_ckv", "b_ckv", "w_cout", "b_cout", "mlp_ln_g",
                                  "mlp_ln_b", "w_mlp1", "b_mlp1", "w_mlp2", "b_mlp2")]


class Weights(C.Structure):
    _fields_ = [("dims", Dims), ("conv1_w", vp), ("conv1_b", vp), ("conv2_w", vp), ("conv2_b", vp), ("enc_pos", vp),
                ("ln_post_g", vp), ("ln_post_b", vp), ("h_enc_layers", C.POINTER(EncLayer)), ("tok_emb", vp),
                ("dec_pos", vp), ("dec_ln_g", vp), ("dec_ln_b", vp), ("h_dec_layers", C.POINTER(DecLayer))]


class DecodeState(C.Structure):
    _fields_ = [("n_seq", i32), ("tokens", vp), ("tokens_ld", i32), ("n_tokens", vp), ("pos", vp),
                ("sum_logprob", vp), ("finished", vp), ("no_speech", vp), ("k_pages", vp), ("v_pages", vp),
                ("layer_page_stride", i64), ("block_table", vp), ("max_pages", i32), ("page_size", i32),
                ("cross_kv", vp), ("cross_layer_stride", i64), ("cross_slot", vp), ("logits", vp),
                ("logits_aux", vp), ("logits_ld", i32), ("suppress_bits", vp), ("xa", vp), ("xa_slots", i32)]


# name -> (restype, argtypes); mirrors include/b200_whisper.h declaration by declaration
SIGNATURES = {
    "b200w_version": (C.c_char_p, []),
    "b200w_last_error": (C.c_char_p, []),
    "b200w_launch_count": (u64, []),
    "b200w_profile_begin": (i32, []),
    "b200w_profile_end": (i32, [C.c_char_p, sz]),
    "b200w_logmel": (i32, [vp, i32, i64, i64, i64, i32, C.POINTER(LogmelTables), vp, vp, vp]),
    "b200w_logmel_pcm16": (i32, [vp, i32, i64, i64, i64, i32, C.POINTER(LogmelTables), vp, vp, vp]),
    "b200w_logmel_normalized": (i32, [vp, i32, i32, i64, i64, i64, i32, C.POINTER(LogmelTables), vp, vp, vp, vp]),
    "b200w_logmel_finalize": (i32, [vp, vp, i32, i64, vp]),
    "b200w_mel_windows": (i32, [vp, vp, vp, vp, vp, i32, i32, vp, vp]),
    "b200w_gemm_bf16": (i32, [vp, i64, vp, vp, i64, vp, vp, i32, i32, i32, i32, vp]),
    "b200w_absorbed_cross_attention_workspace_bytes": (sz, [i32, i32]),
    "b200w_absorbed_cross_attention": (i32, [vp, i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, sz, vp, vp]),
    "b200w_gemm_bf16_splitk": (i32, [vp, i64, vp, vp, i64, i64, i32, i32, i32, i32, vp]),
    "b200w_gemm_splitk_slices": (i32, [i32, i32]),
    "b200w_residual_layernorm": (i32, [vp, vp, i32, i64, vp, vp, vp, i32, i32, vp, vp]),
    "b200w_decoder_self_attention_splitk": (i32, [vp, i32, i64, vp, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp]),
    "b200w_decoder_cross_attention_splitk": (i32, [vp, i32, i64, vp, i32, i32, vp, i64, i32, vp, vp, vp]),
    "b200w_conv1d_gelu": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, i64, i32, vp]),
    "b200w_layernorm": (i32, [vp, vp, vp, i32, i32, vp, vp, vp]),
    "b200w_encoder_attention": (i32, [vp, i32, i32, i32, vp, vp]),
    "b200w_decoder_self_attention": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp]),
    "b200w_decoder_cross_attention": (i32, [vp, i32, i32, i32, vp, i64, i32, vp, vp, vp]),
    "b200w_embed": (i32, [vp, i32, vp, i32, i32, vp, vp, i32, i32, vp, vp]),
    "b200w_filter_argmax": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(FilterParams), vp]),
    "b200w_no_speech_prob": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "b200w_detect_language": (i32, [vp, i32, i32, i32, i32, vp, vp, vp]),
    "b200w_model_create": (i32, [C.POINTER(Weights), C.POINTER(vp)]),
    "b200w_model_destroy": (None, [vp]),
    "b200w_encoder_workspace_bytes": (sz, [vp, i32]),
    "b200w_encoder_forward": (i32, [vp, vp, i32, vp, sz, vp, vp, i32, vp]),
    "b200w_cross_kv": (i32, [vp, vp, i32, vp, i64, i32, vp]),
    "b200w_decoder_workspace_bytes": (sz, [vp, i32, i32]),
    "b200w_decoder_step": (i32, [vp, C.POINTER(DecodeState), i32, i32, i32, C.POINTER(FilterParams), vp, sz, vp]),
    "b200w_decoder_forward_full": (i32, [vp, C.POINTER(DecodeState), i32, vp, sz, vp, vp, i32, vp]),
    "b200w_alignment_matrix": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, i32, i32, vp, vp, vp]),
    "b200w_dtw": (i32, [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp]),
}


def _try_build() -> None:
    import importlib.util

    spec = importlib.util.spec_from_file_location("_b200w_build", os.path.join(_HERE, "csrc", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


def load() -> C.CDLL:
    """Load (building in-tree first if necessary) the CUDA library.  Raises if that is impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            try:
                _try_build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    f"libb200whisper.so is missing at {LIB_PATH} and could not be built ({e}). "
                    "Run `python whisper-mlx_b200/csrc/build.py`; there is no CPU fallback.") from e
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(status: int) -> None:
    if status != 0:
        raise RuntimeError(f"libb200whisper error {status}: {load().b200w_last_error().decode()}")


class kernel_profile:
    """Context manager: CUDA-event timing of every eager kernel launch of the library inside the block.

        with kernel_profile() as prof: ...
        prof.result -> {"kernel": {"launches": n, "total_ms": t}}
    """

    def __enter__(self):
        check(load().b200w_profile_begin())
        self.result = {}
        return self

    def __exit__(self, *exc):
        import json

        buf = C.create_string_buffer(1 << 16)
        n = load().b200w_profile_end(buf, len(buf))
        if n < 0:
            check(n)
        self.result = json.loads(buf.value.decode() or "{}")
        return False


def ptr(t) -> vp:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return vp(0) if t is None else vp(t.data_ptr())


def stream() -> vp:
    import torch

    return vp(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device (got {t.device}); there is no CPU path")
    if not t.is_contiguous():
        raise RuntimeError(f"{what} must be contiguous")
