"""GPU parity cases: the CUDA path (through the C ABI / the package) against the CPU oracle.

Each case is a plain function returning a dict of measured errors; it raises AssertionError on a
parity failure.  tests/test_gpu_parity.py wraps them for pytest (`-m gpu`); tools/gpu_check.py runs each
one in its own subprocess with a timeout so a faulting kernel cannot hide the results of the others.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np
import torch

os.environ.setdefault("B200W_ALLOW_SURROGATE", "1")  # random-init weights, no vocabulary file: ids are what is compared

from tools import synth  # noqa: E402

_CACHE = {}


def _lib():
    from whisper_mlx_b200 import _lib as L

    return L, L.load()


def _model_dir(name: str, seed: int = 0) -> str:
    key = (name, seed)
    if key not in _CACHE:
        d = os.path.join(tempfile.gettempdir(), f"b200w_{name}_{seed}")
        if not os.path.exists(os.path.join(d, "weights.safetensors")):
            synth.write_model(d, name, seed)
        _CACHE[key] = d
    return _CACHE[key]


def _oracle(name: str, seed: int = 0):
    from oracle import model as M

    cfg, w = synth.load_weights_f32(_model_dir(name, seed))
    return M.ModelDimensions(**cfg), w


def _product(name: str, seed: int = 0):
    from whisper_mlx_b200.load_models import load_model

    key = ("prod", name, seed)
    if key not in _CACHE:
        _CACHE[key] = load_model(_model_dir(name, seed))
    return _CACHE[key]


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------ K1 log-mel
# Tolerance named by BASELINE.json north_star: 1e-4 relative.  Outputs live in about [-1, 2]; "relative" is
# taken against max(1, |oracle|), i.e. 1e-4 absolute on the normalised scale.  Signals whose adjacent frames
# differ by > 60 dB (clicks in digital silence) sit at the FP32 FFT noise floor near the -80 dB clamp: there
# the bound is stated separately (1e-3) and the fraction of elements above 1e-4 is reported.
LOGMEL_TOL = 1e-4


def _logmel_err(audio: np.ndarray, n_mels: int, padding: int = 0):
    from oracle import audio as OA
    from whisper_mlx_b200.audio import log_mel_spectrogram

    ref = OA.log_mel_spectrogram(audio, n_mels, padding)
    got = log_mel_spectrogram(audio, n_mels=n_mels, padding=padding).cpu().numpy()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    return float(err.max()), float((err > LOGMEL_TOL).mean())


def case_logmel_noise():
    out = {}
    for n_mels in (80, 128):
        e, frac = _logmel_err(synth.white_noise(480000, 0), n_mels)
        out[f"noise_{n_mels}"] = e
        assert e <= LOGMEL_TOL, (n_mels, e)
    return out


def case_logmel_kinds():
    out = {}
    for kind in ("tones", "speech", "clip", "click"):
        for n_mels in (80, 128):
            e, frac = _logmel_err(synth.make_audio(kind, 160000, 3), n_mels)
            out[f"{kind}_{n_mels}"] = (e, frac)
            tol = LOGMEL_TOL if kind in ("speech", "clip") else 1e-3
            assert e <= tol and frac <= 0.01, (kind, n_mels, e, frac)
    return out


def case_logmel_shapes():
    """Ragged lengths, the `padding` argument, short audio, batched per-row maxima."""
    from oracle import audio as OA
    from whisper_mlx_b200.audio import log_mel_spectrogram

    out = {}
    for n, pad in ((480000, 480000), (16000 * 7 + 123, 0), (4001, 0), (1000, 480000), (160 * 33, 160 * 5), (401, 0),
                   (150, 480000), (1, 1600)):  # shorter than the reflect pad: legal once zero-extended
        e, _ = _logmel_err(synth.white_noise(n, n), 80, pad)
        out[f"n{n}_p{pad}"] = e
        assert e <= LOGMEL_TOL, (n, pad, e)
    x = np.stack([synth.white_noise(48000, 1), 0.01 * synth.white_noise(48000, 2), synth.make_audio("tones", 48000, 3)])
    got = log_mel_spectrogram(x, n_mels=128).cpu().numpy()
    for i in range(3):
        ref = OA.log_mel_spectrogram(x[i], 128)
        e = float(np.abs(got[i] - ref).max())
        out[f"batched_{i}"] = e
        assert e <= (1e-3 if i == 2 else LOGMEL_TOL), (i, e)
    return out


def case_logmel_pcm16():
    """int16 PCM ingestion (the s16le stream load_audio decodes): bit-identical to the f32 path on pcm / 32768, for
    interior tiles (vector loads), edge tiles (reflect, ragged length, padding) and a batch with a row stride."""
    from whisper_mlx_b200.audio import log_mel_spectrogram

    out = {}
    for n, pad, n_mels in ((480000, 0, 80), (16000 * 7 + 123, 480000, 128), (401, 0, 80)):
        pcm = np.clip(np.round(synth.make_audio("speech", n, n) * 32768.0), -32768, 32767).astype(np.int16)
        a = log_mel_spectrogram(pcm, n_mels=n_mels, padding=pad)
        b = log_mel_spectrogram(pcm.astype(np.float32) / 32768.0, n_mels=n_mels, padding=pad)
        out[f"n{n}_p{pad}_{n_mels}"] = float((a - b).abs().max().item())
        assert torch.equal(a, b), (n, pad, n_mels)
    pcm = np.clip(np.round(np.stack([synth.white_noise(48000, 1), synth.make_audio("tones", 48000, 3)]) * 32768.0), -32768, 32767).astype(np.int16)
    a = log_mel_spectrogram(torch.from_numpy(pcm).cuda(), n_mels=128)
    b = log_mel_spectrogram(pcm.astype(np.float32) / 32768.0, n_mels=128)
    assert torch.equal(a, b)
    out["batched"] = 0.0
    return out


# ------------------------------------------------------------------------------------------ K5 GEMM
def _gemm_ref(a, w, bias, gelu, resid):
    y = a.float() @ w.float().T
    if bias is not None:
        y = y + bias
    if gelu:
        y = torch.nn.functional.gelu(y)
    if resid is not None:
        y = y + resid
    return y


def _run_gemm(M, N, K, gelu=False, out_f32=False, bias=True, resid=False, seed=0):
    L, lib = _lib()
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = _bf16(torch.randn(M, K, generator=g)).cuda()
    w = _bf16(torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    r = torch.randn(M, N, generator=g).cuda() if resid else None
    out = torch.empty((M, N), dtype=torch.float32 if out_f32 else torch.bfloat16, device="cuda")
    flags = (1 if gelu else 0) | (2 if out_f32 else 0)
    L.check(lib.b200w_gemm_bf16(L.ptr(a), K, L.ptr(w), L.ptr(out), N, L.ptr(b), L.ptr(r), M, N, K, flags, L.stream()))
    torch.cuda.synchronize()
    ref = _gemm_ref(a, w, b, gelu, r)
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-3 if out_f32 else 1.6e-2) * max(1.0, scale)  # fp32 accumulate; bf16 output rounding = 2^-8 relative
    assert err <= tol, (M, N, K, gelu, out_f32, err, tol)
    return err


def case_gemm():
    out = {}
    shapes = [  # (M, N, K): tile edges, every BLOCK_N variant, K tails
        (128, 128, 64), (300, 384, 384), (1500, 1152, 384), (3000, 1280, 1280), (257, 256, 5120), (129, 1536, 384),
        (4500, 1280, 1280), (64 * 148 + 5, 3840, 1280), (5, 1280, 1280), (120, 5120, 1280), (7, 384, 1536), (3000, 384, 240),
    ]
    for (M, N, K) in shapes:
        out[f"{M}x{N}x{K}"] = _run_gemm(M, N, K)
    out["gelu"] = _run_gemm(1000, 1536, 384, gelu=True)
    out["f32_resid"] = _run_gemm(1000, 384, 1536, out_f32=True, resid=True)
    out["f32_nobias"] = _run_gemm(33, 768, 768, out_f32=True, bias=False)
    return out


def case_gemm_vocab():
    """Tied logits GEMM shape: N = 51866 (not a tile multiple), strided A rows, f32 output with padded ld."""
    from whisper_mlx_b200.whisper import Whisper  # noqa: F401

    m = _product("tiny")
    dims, w = _oracle("tiny")
    E = w["decoder.token_embedding.weight"]
    g = torch.Generator().manual_seed(1)
    h = _bf16(torch.randn(9, dims.n_text_state, generator=g))
    L, lib = _lib()
    ld = m.logits_ld
    out = torch.zeros((9, ld), dtype=torch.float32, device="cuda")
    hc = h.cuda()
    L.check(lib.b200w_gemm_bf16(L.ptr(hc), dims.n_text_state, L.ptr(m.token_embedding), L.ptr(out), ld, None, None, 9,
                                dims.n_vocab, dims.n_text_state, 2, L.stream()))
    torch.cuda.synchronize()
    ref = h.float() @ E.T
    err = (out[:, : dims.n_vocab].cpu() - ref).abs().max().item()
    assert err <= 2e-3, err
    return {"err": err}


def case_conv():
    L, lib = _lib()
    out = {}
    g = torch.Generator().manual_seed(0)
    for (B, T, cin, cout, stride) in ((2, 3000, 80, 384, 1), (3, 3000, 128, 256, 1), (2, 3000, 384, 384, 2)):
        x = _bf16(torch.randn(B, T, cin, generator=g))
        wt = _bf16(torch.randn(cout, 3, cin, generator=g) / (3 * cin) ** 0.5)
        b = torch.randn(cout, generator=g)
        xp = torch.zeros(B, T + 2, cin, dtype=torch.bfloat16)
        xp[:, 1:-1] = x
        t_out = T // stride
        pos = torch.randn(t_out, cout, generator=g) if stride == 2 else None
        o = torch.empty((B * t_out, cout), dtype=torch.float32, device="cuda")
        xpc, wc, bc = xp.cuda(), wt.reshape(cout, 3 * cin).contiguous().cuda(), b.cuda()
        posc = pos.cuda() if pos is not None else None
        L.check(lib.b200w_conv1d_gelu(L.ptr(xpc), L.ptr(wc), L.ptr(bc), B, T, cin, cout, stride, L.ptr(posc), L.ptr(o),
                                      cout, 1, L.stream()))
        torch.cuda.synchronize()
        ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), wt.float().permute(0, 2, 1), b, stride=stride,
                                         padding=1).transpose(1, 2)
        ref = torch.nn.functional.gelu(ref)
        if pos is not None:
            ref = ref + pos
        err = (o.cpu().view(B, t_out, cout) - ref).abs().max().item()
        out[f"{B}x{T}x{cin}->{cout}s{stride}"] = err
        assert err <= 3e-3, err
    return out


# ------------------------------------------------------------------------------------------ K4 / K10
def case_layernorm_embed():
    L, lib = _lib()
    out = {}
    g = torch.Generator().manual_seed(0)
    for d in (384, 768, 1280):
        x = torch.randn(777, d, generator=g) * 3 + 1
        gm, bt = torch.randn(d, generator=g), torch.randn(d, generator=g)
        xc, gc, bc = x.cuda(), gm.cuda(), bt.cuda()  # keep references: a temporary's memory may be reused
        ob = torch.empty((777, d), dtype=torch.bfloat16, device="cuda")
        of = torch.empty((777, d), dtype=torch.float32, device="cuda")
        L.check(lib.b200w_layernorm(L.ptr(xc), L.ptr(gc), L.ptr(bc), 777, d, L.ptr(ob), L.ptr(of), L.stream()))
        torch.cuda.synchronize()
        ref = torch.nn.functional.layer_norm(x, (d,), gm, bt, eps=1e-5)
        e32 = (of.cpu() - ref).abs().max().item()
        e16 = (ob.cpu().float() - ref).abs().max().item()
        out[f"ln{d}"] = (e32, e16)
        assert e32 <= 2e-5 * max(1, ref.abs().max().item()) and e16 <= 2 ** -7 * max(1, ref.abs().max().item())
    # embedding
    V, d, n_ctx = 1000, 384, 448
    te, pe = _bf16(torch.randn(V, d, generator=g)), _bf16(torch.randn(n_ctx, d, generator=g))
    toks = torch.randint(0, V, (5, 456), generator=g, dtype=torch.int32)
    pos = torch.tensor([0, 3, 7, 100, 440], dtype=torch.int32)
    x = torch.empty((5 * 2, d), dtype=torch.float32, device="cuda")
    toks_c, pos_c, te_c, pe_c = toks.cuda(), pos.cuda(), te.cuda(), pe.cuda()
    L.check(lib.b200w_embed(L.ptr(toks_c), 456, L.ptr(pos_c), 5, 2, L.ptr(te_c), L.ptr(pe_c), d, n_ctx,
                            L.ptr(x), L.stream()))
    torch.cuda.synchronize()
    ref = torch.stack([te[toks[b, pos[b] + q].long()].float() + pe[(pos[b] + q).long()].float() for b in range(5) for q in range(2)])
    e = (x.cpu() - ref).abs().max().item()
    out["embed"] = e
    assert e == 0.0, e
    return out


# ------------------------------------------------------------------------------------------ K6 / K7 / K8
def _sdpa_ref(q, k, v, n_head, causal_offset=None):
    """q (B, nq, d), k/v (B, nk, d) f32 -> (B, nq, d); bf16 probabilities like the kernels."""
    B, nq, d = q.shape
    hd = d // n_head
    qh = q.view(B, nq, n_head, hd).transpose(1, 2)
    kh = k.view(B, -1, n_head, hd).transpose(1, 2)
    vh = v.view(B, -1, n_head, hd).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * hd ** -0.5
    if causal_offset is not None:
        nk = kh.shape[2]
        mask = torch.arange(nk)[None, :] > (causal_offset + torch.arange(nq))[:, None]
        s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, nq, d)


def case_encoder_attention():
    L, lib = _lib()
    out = {}
    g = torch.Generator().manual_seed(0)
    for (B, T, H) in ((1, 128, 1), (2, 1500, 6), (1, 1500, 20), (3, 333, 2)):
        d = 64 * H
        qkv = _bf16(torch.randn(B, T, 3 * d, generator=g))
        o = torch.empty((B, T, d), dtype=torch.bfloat16, device="cuda")
        qc = qkv.cuda()
        L.check(lib.b200w_encoder_attention(L.ptr(qc), B, T, H, L.ptr(o), L.stream()))
        torch.cuda.synchronize()
        f = qkv.float()
        ref = _sdpa_ref(f[..., :d], f[..., d: 2 * d], f[..., 2 * d:], H)
        err = (o.cpu().float() - ref).abs().max().item()
        out[f"{B}x{T}x{H}"] = err
        assert err <= 2e-2, (B, T, H, err)
    return out


def case_decoder_attention():
    L, lib = _lib()
    out = {}
    g = torch.Generator().manual_seed(0)
    H, d, ps = 6, 384, 16
    # ---- self attention over pages: sequences at different positions (one, two and four key sweeps), n_q in {1, 3}
    for n_q in (1, 3):
        B, max_pages = 6, 28
        pos = torch.tensor([0, 5, 37, 140, 227, 445], dtype=torch.int32)
        n_pages = B * max_pages
        perm = torch.randperm(n_pages, generator=g).to(torch.int32)  # scattered page table
        bt = perm.view(B, max_pages).contiguous()
        kp = torch.zeros(n_pages, ps, d, dtype=torch.bfloat16)
        vp = torch.zeros_like(kp)
        hist_k = _bf16(torch.randn(B, 448, d, generator=g))
        hist_v = _bf16(torch.randn(B, 448, d, generator=g))
        for b in range(B):
            for j in range(int(pos[b])):
                pg = int(bt[b, j // ps])
                kp[pg, j % ps] = hist_k[b, j]
                vp[pg, j % ps] = hist_v[b, j]
        qkv = _bf16(torch.randn(B, n_q, 3 * d, generator=g))
        o = torch.empty((B, n_q, d), dtype=torch.bfloat16, device="cuda")
        kpc, vpc, qkv_c, pos_c, bt_c = kp.cuda(), vp.cuda(), qkv.cuda(), pos.cuda(), bt.cuda()
        L.check(lib.b200w_decoder_self_attention(L.ptr(qkv_c), B, n_q, H, L.ptr(pos_c), L.ptr(kpc), L.ptr(vpc),
                                                 L.ptr(bt_c), max_pages, ps, L.ptr(o), L.stream()))
        torch.cuda.synchronize()
        errs = []
        for b in range(B):
            p0 = int(pos[b])
            k = torch.cat([hist_k[b, :p0], qkv[b, :, d: 2 * d]], 0).float()[None]
            v = torch.cat([hist_v[b, :p0], qkv[b, :, 2 * d:]], 0).float()[None]
            ref = _sdpa_ref(qkv[b, :, :d].float()[None], k, v, H, causal_offset=p0)
            errs.append((o[b].cpu().float() - ref[0]).abs().max().item())
            # appended rows landed in the right pages
            for qi in range(n_q):
                j = p0 + qi
                pg = int(bt[b, j // ps])
                assert torch.equal(kpc[pg, j % ps].cpu(), qkv[b, qi, d: 2 * d]) and torch.equal(vpc[pg, j % ps].cpu(), qkv[b, qi, 2 * d:])
        out[f"self_nq{n_q}"] = max(errs)
        assert max(errs) <= 2e-2, errs
    # ---- cross attention with a slot table
    T, n_slots = 1500, 4
    for n_q in (1, 2):
        B = 5
        slot = torch.tensor([2, 0, 3, 3, 1], dtype=torch.int32)
        ckv = _bf16(torch.randn(n_slots, T, 2 * d, generator=g))
        q = _bf16(torch.randn(B, n_q, d, generator=g))
        o = torch.empty((B, n_q, d), dtype=torch.bfloat16, device="cuda")
        q_c, ckv_c, slot_c = q.cuda(), ckv.cuda(), slot.cuda()
        L.check(lib.b200w_decoder_cross_attention(L.ptr(q_c), B, n_q, H, L.ptr(ckv_c), T * 2 * d, T,
                                                  L.ptr(slot_c), L.ptr(o), L.stream()))
        torch.cuda.synchronize()
        kv = ckv[slot.long()].float()
        ref = _sdpa_ref(q.float(), kv[..., :d], kv[..., d:], H)
        err = (o.cpu().float() - ref).abs().max().item()
        out[f"cross_nq{n_q}"] = err
        assert err <= 2e-2, err
    return out


def case_absorbed_cross_attention():
    """K14: the absorbed form (streams xa, both contractions on tcgen05) against fp32 attention over K = xa Wk^T and
    V = xa Wv^T + bv rounded to bf16 (what the K / V cache holds), and against K8 on that cache."""
    L, lib = _lib()
    out = {}
    g = torch.Generator().manual_seed(0)
    for (B, H, T, n_slots, n_fin) in ((5, 8, 1500, 3, 0), (37, 12, 1500, 20, 9), (120, 20, 1500, 120, 0), (150, 20, 333, 40, 3), (200, 16, 1500, 7, 11)):
        d = 64 * H
        xa = _bf16(torch.randn(n_slots, T, d, generator=g)).cuda()
        w = _bf16(torch.randn(2 * d, d, generator=g) / d ** 0.5).cuda()
        bias = torch.cat([torch.zeros(d), torch.randn(d, generator=g)]).cuda()
        q = _bf16(torch.randn(B, d, generator=g)).cuda()
        slot = torch.randint(0, n_slots, (B,), generator=g).to(torch.int32).cuda()
        fin = torch.zeros(B, dtype=torch.int32)
        fin[torch.randperm(B, generator=g)[:n_fin]] = 1
        fin = fin.cuda()
        ws = torch.empty(lib.b200w_absorbed_cross_attention_workspace_bytes(B, H), dtype=torch.uint8, device="cuda")
        o = torch.full((B, d), float("nan"), dtype=torch.bfloat16, device="cuda")
        for _ in range(2):  # twice: the arrival counters must be left ready for the next launch
            L.check(lib.b200w_absorbed_cross_attention(L.ptr(q), B, H, L.ptr(w), L.ptr(bias), L.ptr(xa), n_slots, T, L.ptr(slot),
                                                       L.ptr(fin), L.ptr(ws), ws.numel(), L.ptr(o), L.stream()))
        torch.cuda.synchronize()
        kv = (xa.float() @ w.float().T + bias).to(torch.bfloat16)            # (n_slots, T, 2d): the K | V cache
        o8 = torch.empty((B, 1, d), dtype=torch.bfloat16, device="cuda")
        L.check(lib.b200w_decoder_cross_attention(L.ptr(q), B, 1, H, L.ptr(kv), T * 2 * d, T, L.ptr(slot), L.ptr(o8), L.stream()))
        torch.cuda.synchronize()
        kvs = kv[slot.long()].float().cpu()
        ref = _sdpa_ref(q.cpu().float()[:, None], kvs[..., :d], kvs[..., d:], H)[:, 0]
        live = (fin == 0).cpu()
        err = (o.cpu().float() - ref)[live].abs().max().item()
        err8 = (o8[:, 0].cpu().float() - ref)[live].abs().max().item()
        out[f"{B}x{H}x{T}"] = {"absorbed": err, "k8": err8}
        assert err <= 2e-2, (B, H, T, err, err8)
    return out


def case_splitk_decode_ops():
    """Split-K decode GEMM + the consumers that fold the partial slabs (residual+LN, self / cross attention)."""
    L, lib = _lib()
    out = {}
    g = torch.Generator().manual_seed(0)
    H, d, ps, M = 6, 384, 16, 120
    # ---- GEMM partial slabs sum to the product
    for (N, K, split) in ((384, 384, 3), (1280, 1280, 7), (1280, 5120, 7), (3840, 1280, 2), (384, 1536, 8)):
        a = _bf16(torch.randn(M, K, generator=g)).cuda()
        w = _bf16(torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        n_sl = lib.b200w_gemm_splitk_slices(K, split)
        part = torch.full((n_sl, 128, N), float("nan"), dtype=torch.float32, device="cuda")
        L.check(lib.b200w_gemm_bf16_splitk(L.ptr(a), K, L.ptr(w), L.ptr(part), N, 128 * N, M, N, K, split, L.stream()))
        torch.cuda.synchronize()
        ref = a.float() @ w.float().T
        err = (part[:, :M].sum(0) - ref).abs().max().item()
        out[f"gemm_{N}x{K}s{n_sl}"] = err
        assert err <= 2e-3 * max(1.0, ref.abs().max().item()), (N, K, err)
    # ---- residual + LayerNorm
    n_sl = 5
    x = torch.randn(M, d, generator=g).cuda()
    part = (torch.randn(n_sl, 128, d, generator=g) * 0.3).cuda()
    bias, gm, bt = torch.randn(d, generator=g).cuda(), torch.randn(d, generator=g).cuda(), torch.randn(d, generator=g).cuda()
    x_ref = x + bias + part[:, :M].sum(0)
    h_ref = torch.nn.functional.layer_norm(x_ref, (d,), gm, bt, eps=1e-5)
    xo = x.clone()
    h = torch.empty((M, d), dtype=torch.bfloat16, device="cuda")
    L.check(lib.b200w_residual_layernorm(L.ptr(xo), L.ptr(part), n_sl, 128 * d, L.ptr(bias), L.ptr(gm), L.ptr(bt), M, d, L.ptr(h), L.stream()))
    torch.cuda.synchronize()
    out["resid_x"] = (xo - x_ref).abs().max().item()
    out["resid_ln"] = (h.float() - h_ref).abs().max().item()
    assert out["resid_x"] <= 1e-5 and out["resid_ln"] <= 2 ** -7 * max(1.0, h_ref.abs().max().item())
    x2 = x.clone()
    L.check(lib.b200w_residual_layernorm(L.ptr(x2), None, 0, 0, None, L.ptr(gm), L.ptr(bt), M, d, L.ptr(h), L.stream()))
    torch.cuda.synchronize()
    assert torch.equal(x2, x)
    assert (h.float() - torch.nn.functional.layer_norm(x, (d,), gm, bt, eps=1e-5)).abs().max().item() <= 2 ** -7 * 8
    # ---- attention consumers: partial slabs + bias must behave exactly like the bf16 activation they sum to
    B, max_pages, n_sl = 5, 8, 3
    pos = torch.tensor([0, 5, 37, 16, 100], dtype=torch.int32).cuda()
    bt_tab = torch.randperm(B * max_pages, generator=g).to(torch.int32).view(B, max_pages).contiguous().cuda()
    kp = _bf16(torch.randn(B * max_pages, ps, d, generator=g)).cuda()
    vp = _bf16(torch.randn(B * max_pages, ps, d, generator=g)).cuda()
    bias3 = torch.randn(3 * d, generator=g).cuda()
    part3 = torch.randn(n_sl, 128, 3 * d, generator=g).cuda()
    qkv = _bf16(part3[:, :B].sum(0) + bias3).view(B, 1, 3 * d).contiguous()
    o_ref = torch.empty((B, 1, d), dtype=torch.bfloat16, device="cuda")
    o_got = torch.empty_like(o_ref)
    kp1, vp1, kp2, vp2 = kp.clone(), vp.clone(), kp.clone(), vp.clone()
    L.check(lib.b200w_decoder_self_attention(L.ptr(qkv), B, 1, H, L.ptr(pos), L.ptr(kp1), L.ptr(vp1), L.ptr(bt_tab), max_pages,
                                             ps, L.ptr(o_ref), L.stream()))
    L.check(lib.b200w_decoder_self_attention_splitk(L.ptr(part3), n_sl, 128 * 3 * d, L.ptr(bias3), B, H, L.ptr(pos), L.ptr(kp2),
                                                    L.ptr(vp2), L.ptr(bt_tab), max_pages, ps, L.ptr(o_got), L.stream()))
    torch.cuda.synchronize()
    out["self_splitk"] = (o_got.float() - o_ref.float()).abs().max().item()
    assert torch.equal(kp1, kp2) and torch.equal(vp1, vp2), "appended K/V rows differ"
    assert out["self_splitk"] <= 1e-2
    T, n_slots = 1500, 3
    slot = torch.tensor([2, 0, 1, 1, 0], dtype=torch.int32).cuda()
    ckv = _bf16(torch.randn(n_slots, T, 2 * d, generator=g)).cuda()
    bias1 = torch.randn(d, generator=g).cuda()
    part1 = torch.randn(n_sl, 128, d, generator=g).cuda()
    q = _bf16(part1[:, :B].sum(0) + bias1).view(B, 1, d).contiguous()
    L.check(lib.b200w_decoder_cross_attention(L.ptr(q), B, 1, H, L.ptr(ckv), T * 2 * d, T, L.ptr(slot), L.ptr(o_ref), L.stream()))
    L.check(lib.b200w_decoder_cross_attention_splitk(L.ptr(part1), n_sl, 128 * d, L.ptr(bias1), B, H, L.ptr(ckv), T * 2 * d, T,
                                                     L.ptr(slot), L.ptr(o_got), L.stream()))
    torch.cuda.synchronize()
    out["cross_splitk"] = (o_got.float() - o_ref.float()).abs().max().item()
    assert out["cross_splitk"] <= 1e-2
    return out


def case_ring_attention_edges():
    """K8r / K7r (decode attention streamed through cp.async rings) at the edges of their index arithmetic: key counts
    around multiples of the 32-key (cross) / 16-key (self) iteration, the shortest length the ring form takes, the
    longest the kernels accept, page sizes 8 / 16 / 32, scattered slots and pages -- against torch fp32 attention and
    against the register-staged K8 / K7 (B200W_CROSS_STREAM=0 / B200W_SELF_STREAM=0), which must agree bit for bit."""
    L, lib = _lib()
    out = {}
    g = torch.Generator().manual_seed(11)
    H, d = 6, 384
    prev = {k: os.environ.get(k) for k in ("B200W_CROSS_STREAM", "B200W_SELF_STREAM", "B200W_CROSS_PERSIST")}
    try:
        # ---- cross attention: more (sequence, head) units than the key-split form takes, so the full-batch forms run
        B, n_slots = 20, 24
        for T in (256, 257, 300, 1499, 1500, 1536):
            slot = torch.randperm(n_slots, generator=g)[:B].to(torch.int32)
            ckv = _bf16(torch.randn(n_slots, T, 2 * d, generator=g))
            q = _bf16(torch.randn(B, 1, d, generator=g))
            q_c, ckv_c, slot_c = q.cuda(), ckv.cuda(), slot.cuda()
            got = {}
            for form in ("1", "0", "p"):  # K8r, K8, and the opt-in persistent K8p (takes T >= 512)
                os.environ["B200W_CROSS_STREAM"] = "0" if form == "0" else "1"
                os.environ["B200W_CROSS_PERSIST"] = "1" if form == "p" else "0"
                o = torch.full((B, 1, d), float("nan"), dtype=torch.bfloat16, device="cuda")
                L.check(lib.b200w_decoder_cross_attention(L.ptr(q_c), B, 1, H, L.ptr(ckv_c), T * 2 * d, T, L.ptr(slot_c), L.ptr(o), L.stream()))
                torch.cuda.synchronize()
                got[form] = o.cpu()
            os.environ["B200W_CROSS_PERSIST"] = "0"
            kv = ckv[slot.long()].float()
            ref = _sdpa_ref(q.float(), kv[..., :d], kv[..., d:], H)
            err = (got["1"].float() - ref).abs().max().item()
            out[f"cross_T{T}"] = err
            assert err <= 2e-2, (T, err)
            assert torch.equal(got["1"], got["0"]), f"K8r differs from K8 at T = {T}"
            # K8p walks 64 keys per iteration: its partial sums differ in order from K8's, not in value beyond rounding
            assert (got["p"].float() - got["1"].float()).abs().max().item() <= 4e-3, f"K8p differs from K8r at T = {T}"
        # ---- self attention from split-K slabs: positions around the 16-key iterations and the page boundaries
        n_sl = 2
        for ps in (8, 16, 32):
            max_pages = 448 // ps
            pos_l = [0, 1, 15, 16, 17, 31, 32, 127, 128, 129, 255, 300, 446, 447]
            B = len(pos_l)
            pos = torch.tensor(pos_l, dtype=torch.int32).cuda()
            bt_tab = torch.randperm(B * max_pages, generator=g).to(torch.int32).view(B, max_pages).contiguous().cuda()
            kp = _bf16(torch.randn(B * max_pages, ps, d, generator=g)).cuda()
            vp = _bf16(torch.randn(B * max_pages, ps, d, generator=g)).cuda()
            bias3 = torch.randn(3 * d, generator=g).cuda()
            part3 = torch.randn(n_sl, 128, 3 * d, generator=g).cuda()
            res = {}
            for form in ("1", "0"):
                os.environ["B200W_SELF_STREAM"] = form
                kp1, vp1 = kp.clone(), vp.clone()
                o = torch.full((B, 1, d), float("nan"), dtype=torch.bfloat16, device="cuda")
                L.check(lib.b200w_decoder_self_attention_splitk(L.ptr(part3), n_sl, 128 * 3 * d, L.ptr(bias3), B, H, L.ptr(pos), L.ptr(kp1),
                                                                L.ptr(vp1), L.ptr(bt_tab), max_pages, ps, L.ptr(o), L.stream()))
                torch.cuda.synchronize()
                res[form] = (o.cpu(), kp1.cpu(), vp1.cpu())
            assert torch.equal(res["1"][1], res["0"][1]) and torch.equal(res["1"][2], res["0"][2]), "appended K/V rows differ"
            assert torch.equal(res["1"][0], res["0"][0]), f"K7r differs from K7 at page size {ps}"
            # and against fp32 attention over the gathered rows
            qkv = _bf16((part3[:, :B].sum(0) + bias3).cpu())
            kpc, vpc, btc = res["1"][1], res["1"][2], bt_tab.cpu()
            errs = []
            for b in range(B):
                n_keys = pos_l[b] + 1
                rows = [int(btc[b, j // ps]) * ps + j % ps for j in range(n_keys)]
                k = kpc.view(-1, d)[rows].float()[None]
                v = vpc.view(-1, d)[rows].float()[None]
                ref = _sdpa_ref(qkv[b, :d].float()[None, None], k, v, H, causal_offset=pos_l[b])
                errs.append((res["1"][0][b].float() - ref[0]).abs().max().item())
            out[f"self_ps{ps}"] = max(errs)
            assert max(errs) <= 2e-2, (ps, errs)
    finally:
        for k, v in prev.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out


# ------------------------------------------------------------------------------------------ K9
def case_filter_argmax():
    """Every branch of the suppression / timestamp grammar against the oracle's host-side rules."""
    from oracle import decoding as OD
    from oracle.tokens import TokenIds

    L, lib = _lib()
    out = {}
    for n_vocab in (51865, 51866):
        ids = TokenIds(n_vocab)
        tb, eot = ids.timestamp_begin, ids.eot
        sb = 3
        pre = list(ids.sot_sequence("en"))
        hist = [
            [],                                   # first position: must be a timestamp <= 1.00
            [tb + 10],                            # one timestamp: text or a timestamp >= it
            [tb + 10, 400],                       # ts, text
            [tb + 10, 400, tb + 60],              # text then opening ts: no plain text next
            [tb + 10, 400, tb + 60, tb + 60],     # closed pair: text next, ts floor
            [tb + 10, 400, tb + 60, tb + 60, 500, 600],
            [tb + 0],
            [tb + 10, 400, eot],                  # finished sequence: stays EOT
            [tb + 1500],                          # last timestamp
            [tb + 10, tb + 10],
        ]
        rng = np.random.default_rng(n_vocab)
        B = len(hist)
        ld = ((n_vocab + 127) // 128) * 128
        n = max(len(h) for h in hist) + sb
        for variant in range(6):
            logits = rng.standard_normal((B, ld)).astype(np.float32) * 3
            if variant == 1:
                logits[:, tb:] += 6.0        # timestamp mass dominates
            if variant == 2:
                logits[:, tb:] -= 12.0       # text dominates
            if variant == 3:
                logits[:, eot] += 30.0
            if variant == 4:
                logits[:, :] = np.round(logits)  # many exact ties -> lowest index must win
            if variant == 5:
                logits[:, tb: tb + 40] += 9.0
            for b, h in enumerate(hist):
                # per-row histories have different lengths: run each row as its own call
                toks = np.array([pre + h], dtype=np.int64)
                row = logits[b: b + 1, :n_vocab].copy()
                OD.filter_logits(row, toks, sb, ids, ids.suppress_set())
                sum_lp = np.zeros(1, dtype=np.float32)
                new, _ = OD.greedy_update(toks, row, sum_lp, eot)
                exp_tok, exp_lp = int(new[0, -1]), float(sum_lp[0])

                tok_t = torch.zeros((1, 460), dtype=torch.int32)
                tok_t[0, : toks.shape[1]] = torch.from_numpy(toks[0]).int()
                tok_c = tok_t.cuda()
                ntok = torch.tensor([toks.shape[1]], dtype=torch.int32).cuda()
                pos = torch.zeros(1, dtype=torch.int32).cuda()
                slp = torch.zeros(1, dtype=torch.float32).cuda()
                fin = torch.zeros(1, dtype=torch.int32).cuda()
                bits = np.zeros((n_vocab + 31) // 32, dtype=np.uint32)
                for t in ids.suppress_set():
                    bits[t >> 5] |= np.uint32(1 << (t & 31))
                bits_c = torch.from_numpy(bits.view(np.int32)).cuda()
                # rows of 16-byte-aligned length take the vectorised range form of the kernel, any other length the
                # scalar rule chain: both must make the oracle's choice
                ld_odd = n_vocab + (1 if (n_vocab + 1) % 4 else 2)
                for ld_k in ((ld, ld_odd) if variant in (0, 4) else (ld,)):
                    tok_k, ntok_k, pos_k, slp_k, fin_k = tok_c.clone(), ntok.clone(), pos.clone(), slp.clone(), fin.clone()
                    fp = L.FilterParams(n_vocab=n_vocab, logits_ld=ld_k, sample_begin=sb, eot=eot, blank=220,
                                        no_timestamps=ids.no_timestamps, timestamp_begin=tb, no_speech=ids.no_speech,
                                        max_initial_timestamp_index=50, apply_timestamp_rules=1, suppress_blank=1,
                                        tokens_ld=460, temperature=0.0, seed=0)
                    row_k = np.zeros((1, ld_k), dtype=np.float32)
                    row_k[0, :n_vocab] = logits[b, :n_vocab]
                    lg = torch.from_numpy(row_k).cuda()
                    L.check(lib.b200w_filter_argmax(L.ptr(lg), L.ptr(bits_c), L.ptr(tok_k), L.ptr(ntok_k), L.ptr(pos_k), L.ptr(slp_k),
                                                    L.ptr(fin_k), 1, C.byref(fp), L.stream()))
                    torch.cuda.synchronize()
                    got_tok = int(tok_k[0, toks.shape[1]].item())
                    assert got_tok == exp_tok, (n_vocab, variant, b, ld_k, got_tok, exp_tok)
                    assert int(ntok_k.item()) == toks.shape[1] + 1 and int(pos_k.item()) == toks.shape[1]
                    assert int(fin_k.item()) == int(exp_tok == eot)
                    assert abs(float(slp_k.item()) - exp_lp) <= 2e-4 * max(1.0, abs(exp_lp)), (variant, b, ld_k, float(slp_k.item()), exp_lp)
        out[f"v{n_vocab}"] = "ok"
    return out


# ------------------------------------------------------------------------------------------ encoder / decoder
# bf16 tolerance (stated): activations are stored in bf16 (8 significant bits, 2^-9 relative rounding);
# against the fp32 oracle the encoder states (unit scale after ln_post) must agree to 5e-2 max-abs and
# 6e-3 mean-abs; against the oracle run with the same 16-bit storage policy to 3e-2 / 3e-3.
def case_sampling_distribution():
    """T > 0 (the reference's `categorical(logits / T)`, used by the temperature fallback): the on-device Gumbel-max
    sampler is checked against the distribution it has to draw from -- 8192 sequences with the same filtered logits,
    empirical token frequencies vs softmax(filtered / T) within 5 sigma, no forbidden token ever drawn, and the
    log-probability bookkeeping (log_softmax of the UNTEMPERED filtered logits at the drawn token, as in
    GreedyDecoder.update) against the oracle's."""
    from oracle import decoding as OD
    from oracle.tokens import TokenIds

    L, lib = _lib()
    out = {}
    n_vocab = 51866
    ids = TokenIds(n_vocab)
    tb, eot, sb = ids.timestamp_begin, ids.eot, 3
    ld = ((n_vocab + 127) // 128) * 128
    pre = list(ids.sot_sequence("en"))
    B = 8192
    for (hist, T) in (([tb + 10, 400], 0.6), ([tb + 10, 400, tb + 60, tb + 60, 500], 1.0)):
        rng = np.random.default_rng(int(T * 10))
        row = rng.standard_normal(ld).astype(np.float32)
        hot_text = rng.choice(np.arange(1000, 40000), size=6, replace=False)
        row[hot_text] += np.array([9.0, 8.5, 8.0, 7.0, 6.0, 5.0], dtype=np.float32)
        row[tb + 70: tb + 73] += np.array([8.0, 7.0, 6.0], dtype=np.float32)
        row[eot] += 6.5
        toks = np.array([pre + hist], dtype=np.int64)
        filt = row[None, :n_vocab].copy()
        OD.filter_logits(filt, toks, sb, ids, ids.suppress_set())
        ft = torch.from_numpy(filt[0]).double()
        p_ref = torch.softmax(ft / T, dim=-1)
        lp_ref = torch.log_softmax(ft, dim=-1)

        n = toks.shape[1]
        tok_c = torch.zeros((B, 460), dtype=torch.int32)
        tok_c[:, :n] = torch.from_numpy(toks[0]).int()
        tok_c = tok_c.cuda()
        ntok = torch.full((B,), n, dtype=torch.int32).cuda()
        pos = torch.zeros(B, dtype=torch.int32).cuda()
        slp = torch.zeros(B, dtype=torch.float32).cuda()
        fin = torch.zeros(B, dtype=torch.int32).cuda()
        bits = np.zeros((n_vocab + 31) // 32, dtype=np.uint32)
        for t in ids.suppress_set():
            bits[t >> 5] |= np.uint32(1 << (t & 31))
        bits_c = torch.from_numpy(bits.view(np.int32)).cuda()
        fp = L.FilterParams(n_vocab=n_vocab, logits_ld=ld, sample_begin=sb, eot=eot, blank=220,
                            no_timestamps=ids.no_timestamps, timestamp_begin=tb, no_speech=ids.no_speech,
                            max_initial_timestamp_index=50, apply_timestamp_rules=1, suppress_blank=1,
                            tokens_ld=460, temperature=T, seed=1234)
        lg = torch.from_numpy(row).cuda()[None].expand(B, ld).contiguous()
        L.check(lib.b200w_filter_argmax(L.ptr(lg), L.ptr(bits_c), L.ptr(tok_c), L.ptr(ntok), L.ptr(pos), L.ptr(slp),
                                        L.ptr(fin), B, C.byref(fp), L.stream()))
        torch.cuda.synchronize()
        drawn = tok_c[:, n].cpu().long()
        assert bool(torch.isfinite(ft[drawn]).all()), "a token the rules forbid was drawn"
        counts = torch.bincount(drawn, minlength=n_vocab).double()
        support = torch.nonzero(p_ref > 1e-3)[:, 0]
        worst = 0.0
        for v in support.tolist():
            pv = float(p_ref[v])
            sigma = (B * pv * (1 - pv)) ** 0.5
            z = abs(float(counts[v]) - B * pv) / sigma
            worst = max(worst, z)
            assert z <= 5.0, (T, v, float(counts[v]), B * pv, sigma)
        rest = float(counts.sum() - counts[support].sum()) / B
        assert abs(rest - float(1 - p_ref[support].sum())) <= 0.01
        e_lp = (slp.cpu().double() - lp_ref[drawn]).abs().max().item()
        assert e_lp <= 2e-4 * max(1.0, float(lp_ref[drawn].abs().max())), e_lp
        assert torch.equal(fin.cpu().long(), (drawn == eot).long())
        out[f"T{T}"] = {"support": len(support), "worst_sigma": worst, "distinct_drawn": int((counts > 0).sum()), "logprob_err": e_lp}
        assert int((counts > 0).sum()) >= len(support)
    return out


def case_encoder_tiny():
    from oracle import audio as OA, model as OM

    dims, w = _oracle("tiny")
    m = _product("tiny")
    out = {}
    x = np.stack([synth.white_noise(480000, 0), synth.make_audio("speech", 480000, 1)])
    mel = np.stack([OA.log_mel_spectrogram(a, dims.n_mels) for a in x])
    mel_t = torch.from_numpy(mel)
    ref32 = OM.encoder_forward(w, dims, mel_t, policy="fp32")
    ref16 = OM.encoder_forward(w, dims, mel_t, policy="bf16")
    slabs = m._mel_to_slabs(mel_t)
    # conv stem probe (0 transformer blocks)
    _, stem = m.encode_slabs(slabs, stop_after_layers=0)
    stem_ref = OM.encoder_forward(w, dims, mel_t, policy="bf16", return_stem=True)
    e = (stem.cpu() - stem_ref).abs().max().item()
    out["stem"] = e
    assert e <= 2e-2, e
    xa, xa32 = m.encode_slabs(slabs, want_f32=True)
    torch.cuda.synchronize()
    for name, ref, tmax, tmean in (("fp32", ref32, 5e-2, 6e-3), ("bf16", ref16, 3e-2, 3e-3)):
        d = (xa32.cpu() - ref).abs()
        out[name] = (d.max().item(), d.mean().item())
        assert d.max().item() <= tmax and d.mean().item() <= tmean, (name, d.max().item(), d.mean().item())
    assert (xa.float() - xa32).abs().max().item() <= 2 ** -7 * max(1.0, xa32.abs().max().item())
    return out


def case_decoder_tiny():
    """Teacher-forced logits and argmax ids over a token sequence that walks the timestamp grammar."""
    from oracle import audio as OA, model as OM
    from oracle.tokens import TokenIds

    dims, w = _oracle("tiny")
    m = _product("tiny")
    ids = TokenIds(dims.n_vocab)
    tb = ids.timestamp_begin
    x = np.stack([synth.white_noise(480000, 5), synth.make_audio("tones", 480000, 6)])
    mel_t = torch.from_numpy(np.stack([OA.log_mel_spectrogram(a, dims.n_mels) for a in x]))
    xa_ref = OM.encoder_forward(w, dims, mel_t, policy="bf16")
    seq = list(ids.sot_sequence("en")) + [tb + 5, 300, 4000, 17, tb + 80, tb + 80, 900, 901, 902, tb + 200, tb + 200, 12, 13, tb + 400, ids.eot]
    toks = torch.tensor([seq, seq], dtype=torch.long)
    ref32, _ = OM.decoder_forward(w, dims, toks, xa_ref, policy="fp32")
    ref16, _ = OM.decoder_forward(w, dims, toks, xa_ref, policy="bf16")
    # feed the product the oracle's encoder states so the decoder is checked in isolation
    got = m.logits(toks, xa_ref.to(torch.bfloat16).cuda()).cpu()
    out = {}
    # stated bf16 tolerance for logits (std ~1): 6e-2 max-abs vs fp32 oracle, 4e-2 vs the 16-bit-policy oracle
    for name, ref, tol in (("fp32", ref32, 6e-2), ("bf16", ref16, 4e-2)):
        d = (got - ref).abs().max().item()
        out[name] = d
        assert d <= tol, (name, d)
    # argmax identity, margin-aware: identical wherever the oracle's top-1/top-2 margin
    # exceeds twice the observed logit error; report the rest
    top2 = ref16.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    same = got.argmax(-1) == ref16.argmax(-1)
    safe = margin > 2 * out["bf16"]
    out["argmax_equal"] = int(same.sum().item())
    out["positions"] = int(same.numel())
    assert bool(same[safe].all()), "argmax differs at a position with a safe margin"
    assert same.float().mean().item() >= 0.9
    return out


# Free-running parity, margin-aware.  Two implementations of the same greedy decoder diverge for
# good at the first near-tie that rounding flips, so "the same tokens" can only be demanded where the oracle's own
# margin exceeds the numerical error.  Instead of counting an equal prefix, EVERY token the CUDA path sampled is checked
# against the oracle run teacher-forced on the CUDA path's own history:
#   * err  = max |product logits - oracle logits| over all positions of the sequence (measured, and bounded by the
#            stated bf16 tolerance);
#   * at every position the sampled token's filtered oracle logit must lie within 2 * err of the oracle's best
#     (it IS the oracle's argmax whenever the oracle's top-1 / top-2 margin exceeds 2 * err);
#   * positions where it is not the argmax are near-ties and are counted; with none, the free-running oracle
#     sequence must be identical.
LOGIT_TOL_BF16 = 4e-2   # vs the 16-bit-storage-policy oracle, logits of std ~1 (case_decoder_tiny)


def _check_greedy_trajectory(w, dims, xa_ref, product_logits, sampled, n_steps, language="en", logit_tol=LOGIT_TOL_BF16,
                             avg_logprob=None, greedy=True):
    """xa_ref (1500, d) f32 oracle-side encoder states of ONE window; `sampled`: the tokens the CUDA path produced
    (EOT-trimmed); product_logits: callable(tokens (1, n) long) -> (n, V) f32 teacher-forced logits of the CUDA path.
    Returns {"err", "near_ties", "min_margin_at_tie"}; raises AssertionError when a token is not explained by rounding."""
    from oracle import decoding as OD, model as OM
    from oracle.tokens import TokenIds

    ids = TokenIds(dims.n_vocab)
    initial = list(ids.sot_sequence(language))
    sb = len(initial)
    ended = len(sampled) < n_steps  # an EOT was sampled
    seq = initial + [int(t) for t in sampled] + ([ids.eot] if ended else [])
    toks = torch.tensor([seq], dtype=torch.long)
    ref, _ = OM.decoder_forward(w, dims, toks, xa_ref[None], policy="bf16")
    ref = ref[0].float()
    got = product_logits(toks)
    err = float((got - ref).abs().max())
    assert err <= logit_tol, ("teacher-forced logits off", err)
    thr = 2.0 * err
    near, worst, sum_lp = 0, None, 0.0
    for j in range(sb - 1, len(seq) - 1):  # logits at j choose seq[j + 1]
        row = ref[j: j + 1].numpy().copy()
        hist = np.array([seq[: j + 1]], dtype=np.int64)
        OD.filter_logits(row, hist, sb, ids, ids.suppress_set())
        choice = seq[j + 1]
        best = float(row.max())
        assert np.isfinite(row[0, choice]), (j, choice, "the CUDA path sampled a token the rules forbid")
        gap = best - float(row[0, choice])
        assert gap <= thr or not greedy, (j, choice, int(row.argmax()), gap, thr, "token differs beyond rounding")
        if int(row.argmax()) != choice:
            near += 1
            worst = gap if worst is None else max(worst, gap)
        sum_lp += float(torch.log_softmax(torch.from_numpy(row[0]), -1)[choice])
    out = {"err": err, "near_ties": near, "worst_gap": worst, "n": len(seq) - sb}
    if avg_logprob is not None:
        ref_avg = sum_lp / (len(sampled) + 1)
        out["avg_logprob"] = (avg_logprob, ref_avg)
        assert abs(avg_logprob - ref_avg) <= 2 * err + 1e-3, (avg_logprob, ref_avg)
    return out


def _product_logits_fn(m, xa_dev):
    """Teacher-forced logits of the CUDA path for one window whose encoder states are on the device."""
    def fn(toks):
        return m.logits(toks, xa_dev)[0].float().cpu()
    return fn


def case_decode_tiny():
    """Free-running greedy decode (K9 on device, CUDA-graph steps) vs the oracle: every sampled token is the oracle's
    choice up to rounding (see _check_greedy_trajectory); with no near-tie the sequences are identical."""
    from oracle import audio as OA, decoding as OD, model as OM
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask

    dims, w = _oracle("tiny")
    m = _product("tiny")
    x = np.stack([synth.white_noise(480000, 7), synth.make_audio("speech", 480000, 8), synth.make_audio("tones", 480000, 9)])
    mel_t = torch.from_numpy(np.stack([OA.log_mel_spectrogram(a, dims.n_mels) for a in x]))
    xa_ref = OM.encoder_forward(w, dims, mel_t, policy="bf16")
    n_steps = 40
    ref = OD.decode(w, dims, mel_t, language="en", sample_len=n_steps, policy="bf16", audio_features=xa_ref)
    task = DecodingTask(m, DecodingOptions(language="en", sample_len=n_steps))
    xa_dev = xa_ref.to(torch.bfloat16).cuda()
    got = task.run_features(xa_dev)
    out = {}
    for i, (g, r) in enumerate(zip(got, ref)):
        chk = _check_greedy_trajectory(w, dims, xa_ref[i], _product_logits_fn(m, xa_dev[i: i + 1]), g.tokens, n_steps,
                                       avg_logprob=g.avg_logprob)
        out[f"w{i}"] = chk
        assert abs(g.no_speech_prob - r.no_speech_prob) <= 0.05 * max(r.no_speech_prob, 1e-6) + 1e-7
        if chk["near_ties"] == 0:
            assert g.tokens == r.tokens, (i, "no near-tie, yet the free-running sequences differ")
            assert abs(g.avg_logprob - r.avg_logprob) <= 2e-2, (g.avg_logprob, r.avg_logprob)
    return out


def case_large_v3_parity():
    """The named architecture (large-v3: 128 mel, d=1280, 20 heads, 32+32 layers, V=51866) with random-init weights:
    encoder states, teacher-forced logits and argmax ids against the CPU oracle on one 30 s window."""
    from oracle import audio as OA, model as OM
    from oracle.tokens import TokenIds
    from whisper_mlx_b200.whisper import ModelDimensions, Whisper

    dims_d = synth.DIMS["large-v3"]
    weights = dict(synth.random_weights(dims_d, 0, device="cuda"))
    m = Whisper(ModelDimensions(**dims_d), weights)
    w32 = {k: v.detach().to("cpu", torch.float32) for k, v in weights.items()}
    del weights
    dims = OM.ModelDimensions(**dims_d)
    ids = TokenIds(dims.n_vocab)
    tb = ids.timestamp_begin
    torch.set_num_threads(os.cpu_count() or 1)
    # three windows: M = 4500 rows puts the bf16-output projections on the 2-CTA (cta_group::2) GEMM
    mel_t = torch.from_numpy(np.stack([OA.log_mel_spectrogram(synth.make_audio(k, 480000, 4 + i), dims.n_mels)
                                       for i, k in enumerate(("speech", "noise", "tones"))]))
    ref = OM.encoder_forward(w32, dims, mel_t, policy="bf16")
    xa, xa32 = m.encode_slabs(m._mel_to_slabs(mel_t), want_f32=True)
    d = (xa32.cpu() - ref).abs()
    out = {"enc_max": d.max().item(), "enc_mean": d.mean().item(), "enc_ref_absmax": ref.abs().max().item()}
    # 32 layers of bf16 storage: stated tolerance 8e-2 max-abs / 8e-3 mean-abs on unit-scale states
    assert out["enc_max"] <= 8e-2 and out["enc_mean"] <= 8e-3, out
    seq = list(ids.sot_sequence("en")) + [tb + 5, 300, 4000, tb + 80, tb + 80, 900, 901, tb + 200]
    toks = torch.tensor([seq] * 3, dtype=torch.long)
    ref_l, _ = OM.decoder_forward(w32, dims, toks, ref, policy="bf16")
    got = m.logits(toks, ref.to(torch.bfloat16).cuda()).cpu()
    out["logit_max"] = (got - ref_l).abs().max().item()
    assert out["logit_max"] <= 8e-2, out
    top2 = ref_l.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    same = got.argmax(-1) == ref_l.argmax(-1)
    out["argmax_equal"] = f"{int(same.sum())}/{same.numel()}"
    assert bool(same[margin > 2 * out["logit_max"]].all()), "argmax differs at a position with a safe margin"
    return out


def _config_decode_case(name: str, batch: int, n_steps: int):
    """A BASELINE decode configuration at its full batch size: greedy decode of `batch` windows of which only four are
    distinct.  Size-independent properties: every copy decodes exactly like its original (rows of a batch never
    interact), and each of the four originals is a greedy trajectory of the oracle up to rounding -- every token,
    margin-aware (_check_greedy_trajectory); the teacher-forced check runs the same single-token chain path at batch 1."""
    from oracle import decoding as OD, model as OM
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask
    from whisper_mlx_b200.whisper import ModelDimensions, Whisper

    dims_d = synth.DIMS[name]
    weights = dict(synth.random_weights(dims_d, 0, device="cuda"))
    m = Whisper(ModelDimensions(**dims_d), weights)
    w32 = {k: v.detach().to("cpu", torch.float32) for k, v in weights.items() if k.startswith("decoder.")}
    del weights
    dims = OM.ModelDimensions(**dims_d)
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(11)
    base = _bf16(torch.randn(4, dims.n_audio_ctx, dims.n_audio_state, generator=g))
    xa = base[torch.arange(batch) % 4].cuda()
    from whisper_mlx_b200.decoding import total_kernel_launches

    k0 = total_kernel_launches()
    got = DecodingTask(m, DecodingOptions(language="en", sample_len=n_steps)).run_features(xa)
    assert len(got) == batch
    out = {"batch": batch, "kernel_launches": total_kernel_launches() - k0}
    for i in range(4, batch):
        assert got[i].tokens == got[i % 4].tokens and got[i].avg_logprob == got[i % 4].avg_logprob, (i, "copy differs")
        assert got[i].no_speech_prob == got[i % 4].no_speech_prob
    ref = OD.decode(w32, dims, torch.zeros(4, 3000, dims.n_mels), language="en", sample_len=n_steps, policy="bf16",
                    audio_features=base.float())
    for i in range(4):
        chk = _check_greedy_trajectory(w32, dims, base[i].float(), _product_logits_fn(m, xa[i: i + 1]), got[i].tokens,
                                       n_steps, logit_tol=8e-2,  # unit-variance random encoder states, up to 32 layers
                                       avg_logprob=got[i].avg_logprob)
        out[f"w{i}"] = chk
        assert abs(got[i].no_speech_prob - ref[i].no_speech_prob) <= 0.05 * max(ref[i].no_speech_prob, 1e-6) + 1e-7
        if chk["near_ties"] == 0:
            assert got[i].tokens == ref[i].tokens, (i, "no near-tie, yet the free-running sequences differ")
    m.release_sessions()
    return out


def case_config4_large_v3_batch120():
    """BASELINE configs[3] at its own shape: whisper-large-v3 (32 decoder layers, d = 1280), 120 windows decoded
    together -- the benchmarked path: prompt step, then single-token steps replayed from the CUDA graph through the
    single-tile split-K decode chain (K11) and the batch-120 attention kernels."""
    out = _config_decode_case("large-v3", 120, 26)
    assert out["kernel_launches"] >= 25 * 100, out  # 25 graph-replayed steps of >= 100 kernels each did run
    return out


def case_config2_logmel_batch1024():
    """BASELINE configs[1]: the log-mel front-end on 1024 x 30 s windows, 80 and 128 mel bins.  Four distinct signals are
    tiled over the batch: every copy must equal its original bit for bit (tiles never interact, each row has its own
    clamp maximum) and the originals must match the oracle."""
    from oracle import audio as OA
    from whisper_mlx_b200.audio import log_mel_spectrogram

    base = np.stack([synth.white_noise(480000, 0), 0.01 * synth.white_noise(480000, 1), synth.make_audio("speech", 480000, 2),
                     synth.make_audio("tones", 480000, 3)]).astype(np.float32)
    x = torch.from_numpy(base).cuda()[torch.arange(1024, device="cuda") % 4].contiguous()
    out = {}
    for n_mels in (80, 128):
        got = log_mel_spectrogram(x, n_mels=n_mels)
        assert got.shape == (1024, 3000, n_mels)
        assert torch.equal(got[4:], got[:4].repeat(255, 1, 1)), "a copy differs from its original"
        for i in range(4):
            ref = OA.log_mel_spectrogram(base[i], n_mels)
            e = float((np.abs(got[i].cpu().numpy() - ref) / np.maximum(1.0, np.abs(ref))).max())
            out[f"mel{n_mels}_w{i}"] = e
            assert e <= (1e-3 if i == 3 else LOGMEL_TOL), (n_mels, i, e)
        del got
    return out


def case_config3_small_batch64():
    """BASELINE configs[2]: whisper-small greedy decode, batch 64 (split-K chain path: rows <= 128)."""
    return _config_decode_case("small", 64, 24)


def case_config5_turbo_batch256():
    """BASELINE configs[4]: whisper-large-v3-turbo (4-layer decoder), batch 256 (general path: rows > 128)."""
    return _config_decode_case("large-v3-turbo", 256, 16)


def case_small_batch_step():
    """K13 / K13m (one cooperative launch per single-token step: FP32-pipe projections for <= 2 sequences -- the exact
    sequential mode -- and mma.sync projections for 3 .. 7 -- its `best_of` fallback) against the large-batch path (K11
    chains + attention kernels) and the oracle: teacher-forced logits over a token sequence that crosses a KV page
    boundary, batches 1 .. 16 with DIFFERENT windows per row, then free-running greedy decoding."""
    from oracle import model as OM
    from oracle.tokens import TokenIds
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask
    from whisper_mlx_b200.whisper import ModelDimensions, Whisper

    dims_d = synth.DIMS["small"]
    weights = dict(synth.random_weights(dims_d, 0, device="cuda"))
    m = Whisper(ModelDimensions(**dims_d), weights)
    w32 = {k: v.detach().to("cpu", torch.float32) for k, v in weights.items() if k.startswith("decoder.")}
    del weights
    dims = OM.ModelDimensions(**dims_d)
    ids = TokenIds(dims.n_vocab)
    tb = ids.timestamp_begin
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(5)
    xa = _bf16(torch.randn(6, dims.n_audio_ctx, dims.n_audio_state, generator=g) * 0.7)
    seq = list(ids.sot_sequence("en")) + [tb + 5, 300, 4000, 17, tb + 80, tb + 80, 900, 901, 902, tb + 200, tb + 200, 12, 13, 14, 15, 16,
                                          tb + 400, tb + 400, 21]
    out = {}
    prev = os.environ.get("B200W_SMALL")
    prev_mma = os.environ.get("B200W_SMALL_MMA")
    try:
        # (batch, form): K13 = FP32-pipe projections (forced up to 5 sequences), K13m = mma.sync projections (its default
        # range 3 .. 7 and, forced, the two-n-tile kernels up to 16 sequences)
        for B, form in ((1, "K13"), (2, "K13"), (3, "K13"), (5, "K13"), (3, "K13m"), (5, "K13m"), (7, "K13m"), (9, "K13m"), (16, "K13m")):
            toks = torch.tensor([seq] * B, dtype=torch.long)
            toks[:, 6] += torch.arange(B)  # rows differ in their history too
            xa_b = xa[torch.arange(B) % xa.shape[0]]
            os.environ["B200W_SMALL"] = "1"
            os.environ["B200W_SMALL_MMA"] = "all" if form == "K13m" else "0"
            k0 = _lib()[1].b200w_launch_count()
            got = m.logits(toks, xa_b.cuda()).cpu()
            n_small = _lib()[1].b200w_launch_count() - k0
            os.environ["B200W_SMALL"] = "0"
            os.environ["B200W_SMALL_MMA"] = "0"
            k0 = _lib()[1].b200w_launch_count()
            big = m.logits(toks, xa_b.cuda()).cpu()
            n_big = _lib()[1].b200w_launch_count() - k0
            assert n_small < n_big / 4, (n_small, n_big)  # the one-launch path really ran
            ref, _ = OM.decoder_forward(w32, dims, toks, xa_b.float(), policy="bf16")
            e_paths = (got - big).abs().max().item()
            e_ref = (got - ref).abs().max().item()
            e_big = (big - ref).abs().max().item()
            out[f"{form}_B{B}"] = {"small_vs_chain": e_paths, "small_vs_oracle": e_ref, "chain_vs_oracle": e_big, "launches": (n_small, n_big)}
            assert e_ref <= LOGIT_TOL_BF16 * 1.5 and e_paths <= LOGIT_TOL_BF16 * 1.5, out
            top2 = ref.topk(2, dim=-1).values
            safe = (top2[..., 0] - top2[..., 1]) > 2 * e_ref
            assert bool((got.argmax(-1) == ref.argmax(-1))[safe].all())
        os.environ.pop("B200W_SMALL_MMA", None)
        # free-running: graph-replayed one-launch steps, every token a greedy choice of the oracle up to rounding
        os.environ["B200W_SMALL"] = "1"
        n_steps = 40
        res = DecodingTask(m, DecodingOptions(language="en", sample_len=n_steps)).run_features(xa[:2].cuda())
        for i in range(2):
            out[f"free{i}"] = _check_greedy_trajectory(w32, dims, xa[i].float(), _product_logits_fn(m, xa[i: i + 1].cuda()), res[i].tokens,
                                                       n_steps, logit_tol=8e-2, avg_logprob=res[i].avg_logprob)
        # ... and five windows at once: K13m (the default for 3 .. 7 sequences), graph-replayed
        os.environ.pop("B200W_SMALL", None)
        k0 = _lib()[1].b200w_launch_count()
        res5 = DecodingTask(m, DecodingOptions(language="en", sample_len=24)).run_features(xa[:5].cuda())
        assert _lib()[1].b200w_launch_count() - k0 < 24 * 20, "five sequences did not take the one-launch path"
        for i in (0, 4):
            out[f"free_mma{i}"] = _check_greedy_trajectory(w32, dims, xa[i].float(), _product_logits_fn(m, xa[i: i + 1].cuda()), res5[i].tokens,
                                                           24, logit_tol=8e-2, avg_logprob=res5[i].avg_logprob)
    finally:
        if prev is None:
            os.environ.pop("B200W_SMALL", None)
        else:
            os.environ["B200W_SMALL"] = prev
        if prev_mma is None:
            os.environ.pop("B200W_SMALL_MMA", None)
        else:
            os.environ["B200W_SMALL_MMA"] = prev_mma
    m.release_sessions()
    return out


def case_decode_dual_stream():
    """Batches of >= 16 windows decode as two half-batches on two streams; the result must not depend on it."""
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask

    m = _product("tiny")
    dims, _ = _oracle("tiny")
    g = torch.Generator().manual_seed(3)
    # 32 windows: both the whole batch and its halves are above the size at which the cross-attention splits its keys
    # (a different, equally valid summation order; see cross_attention_kv_splits)
    xa = _bf16(torch.randn(32, dims.n_audio_ctx, dims.n_audio_state, generator=g)).cuda()
    outs = []
    for n_streams in (1, 2, 2):
        m.decode_streams = n_streams
        task = DecodingTask(m, DecodingOptions(language="en", sample_len=48))
        res = task.run_features(xa)
        outs.append([(r.tokens, round(r.avg_logprob, 6), round(r.no_speech_prob, 9)) for r in res])
    m.decode_streams = 1
    assert len(outs[0]) == 32
    assert outs[0] == outs[1] == outs[2], "two-stream decoding changed the result"
    return {"tokens_first": outs[0][0][0][:8], "n": len(outs[0])}


def _replay_and_validate(mode, fixed, got, trace, w, dims, audio, m, kw, n_steps):
    """(1) The reference control flow (oracle/transcribe.py), replayed on the tokens the CUDA path decoded window by
    window, must reproduce the product's segments exactly -- all of them: ids, seek, start, end, tokens, text.
    (2) Every decoded window is a greedy trajectory of the oracle up to rounding."""
    from oracle import audio as OA, decoding as OD, model as OM, transcribe as OT

    by_key = {(t["seek"], t["size"]): t for t in trace}
    used = []

    def replay(seek, size, segment_mel, prompt, temperature):
        t = by_key[(seek, size)]  # KeyError: the product never decoded the window the reference flow asks for
        used.append((seek, size))
        return OD.DecodingResult(tokens=list(t["tokens"]), text=t["text"], avg_logprob=t["avg_logprob"],
                                 no_speech_prob=t["no_speech_prob"], temperature=t["temperature"],
                                 compression_ratio=t["compression_ratio"])

    ref = OT.transcribe(w, dims, audio, policy="bf16", fixed_windows=fixed, decode_fn=replay, **kw)
    assert sorted(used) == sorted(by_key), (mode, "windows decoded by the product but not by the reference flow", used)
    assert len(ref["segments"]) == len(got["segments"]), (mode, len(ref["segments"]), len(got["segments"]))
    for gs, rs in zip(got["segments"], ref["segments"]):
        assert set(gs) == set(rs)
        for k in ("id", "seek", "tokens", "text", "temperature"):
            assert gs[k] == rs[k], (mode, gs["id"], k, gs[k], rs[k])
        for k in ("start", "end", "avg_logprob", "no_speech_prob", "compression_ratio"):
            assert abs(gs[k] - rs[k]) < 1e-6, (mode, gs["id"], k, gs[k], rs[k])
    assert got["text"] == ref["text"] and got["language"] == ref["language"]
    # every window against the oracle model on the oracle's own log-mel / encoder
    mel = OA.log_mel_spectrogram(audio, dims.n_mels, padding=OA.N_SAMPLES)
    near = 0
    for (seek, size), t in sorted(by_key.items()):
        seg = torch.from_numpy(OA.pad_or_trim(mel[seek: seek + size], 3000, axis=-2))[None]
        xa_ref = OM.encoder_forward(w, dims, seg, policy="bf16")
        xa_dev = m.embed_audio(seg)
        chk = _check_greedy_trajectory(w, dims, xa_ref[0], _product_logits_fn(m, xa_dev), t["tokens"], n_steps,
                                       logit_tol=6e-2, avg_logprob=t["avg_logprob"])  # encoder + decoder rounding
        near += chk["near_ties"]
    return len(got["segments"]), len(by_key), near


def case_transcribe_micro():
    """End-to-end `transcribe()` (log-mel -> encoder -> decode -> segments) vs the oracle, both modes: ALL segments equal
    the reference control flow replayed on the decoded tokens, and every window's tokens follow the oracle model up to
    rounding; when no near-tie occurred anywhere the free-running oracle transcript must be identical too."""
    from oracle import transcribe as OT
    from whisper_mlx_b200 import transcribe

    dims, w = _oracle("micro")
    m = _product("micro")
    audio = synth.long_audio(75.0, 3)
    out = {}
    n_steps = 24
    kw = dict(temperature=0.0, condition_on_previous_text=False, language="en", sample_len=n_steps)
    for mode, fixed in (("exact", False), ("batched", True)):
        trace = []
        got = transcribe(audio, model=m, window_batch=4 if fixed else 0, window_trace=trace, **kw)
        n_seg, n_win, near = _replay_and_validate(mode, fixed, got, trace, w, dims, audio, m, kw, n_steps)
        free = OT.transcribe(w, dims, audio, policy="bf16", fixed_windows=fixed, **kw)
        n_same = 0
        for gs, rs in zip(got["segments"], free["segments"]):
            if gs["tokens"] == rs["tokens"] and abs(gs["start"] - rs["start"]) < 1e-6 and abs(gs["end"] - rs["end"]) < 1e-6:
                n_same += 1
            else:
                break
        out[mode] = {"segments": n_seg, "windows": n_win, "near_ties": near, "equal_to_free_running_oracle": n_same,
                     "oracle_segments": len(free["segments"])}
        if near == 0:
            assert n_same == len(free["segments"]) == n_seg, (mode, out[mode])
    return out


def case_word_alignment():
    """Word-level timestamps (timing.py): cross-attention probabilities of the alignment heads, the normalise / median /
    head-mean matrix and the DTW path against the oracle, then `transcribe(word_timestamps=True)` end to end."""
    from oracle import audio as OA, model as OM, timing as OT
    from whisper_mlx_b200 import timing as T, transcribe
    from whisper_mlx_b200.tokenizer import get_tokenizer

    out = {}
    dims, w = _oracle("tiny")
    m = _product("tiny")
    tk = get_tokenizer(True, num_languages=m.num_languages, language="en", task="transcribe")
    mel = torch.from_numpy(OA.log_mel_spectrogram(synth.make_audio("speech", 480000, 11), dims.n_mels))[None]
    xa = OM.encoder_forward(w, dims, mel, policy="bf16")
    rng = np.random.default_rng(5)
    text_tokens = [int(t) for t in rng.integers(300, 20000, size=37)]
    heads = np.asarray(m.alignment_heads).reshape(-1, 2)
    for num_frames in (3000, 1724):
        # ---- oracle: scores -> matrix -> path
        tokens = torch.tensor([*tk.sot_sequence, tk.no_timestamps, *text_tokens, tk.eot], dtype=torch.long)
        logits_ref, _, cross_qk = OM.decoder_forward(w, dims, tokens[None], xa, policy="bf16", return_cross_qk=True)
        mat_ref = OT.alignment_matrix(cross_qk, heads, num_frames)
        # ---- GPU: probabilities from K8, K12a/b matrix, K12c path
        n = len(tokens)
        sess = m.decode_session(1, 1, dims.n_text_ctx, slot=7)
        sess.load(_bf16(xa).cuda())
        sess.set_tokens(tokens[None].to(torch.int32))
        first = int(heads[:, 0].min())
        logits, probs = sess.forward_full(n, first)
        torch.cuda.synchronize()
        lerr = (logits.cpu() - logits_ref[0]).abs().max().item()
        p_ref = torch.softmax(torch.stack([cross_qk[l][0] for l in range(first, dims.n_text_layer)]).float(), -1)  # (L', H, n, T)
        perr = (probs[:, 0].permute(0, 2, 1, 3).cpu() - p_ref).abs().max().item()
        assert abs(probs[:, 0].sum(-1).mean().item() - 1.0) < 1e-3
        mat = T.alignment_matrix(m, probs, first, num_frames)
        merr = float(np.abs(mat.cpu().numpy() - mat_ref).max())
        out[f"f{num_frames}"] = {"logits": lerr, "probs": perr, "matrix": merr}
        assert lerr <= 6e-2 and perr <= 2e-4 and merr <= 0.1, out  # z-score units; measured 0.019 / 2.7e-5 / 0.032
        # the matrix stage alone, on the GPU's own probabilities (isolates K12a/b from model rounding): tight
        fake_qk = [None] * first + [torch.log(probs[i, 0].permute(1, 0, 2).cpu().double())[None] for i in range(probs.shape[0])]
        mat_same = OT.alignment_matrix(fake_qk, heads, num_frames)
        m2 = float(np.abs(mat.cpu().numpy() - mat_same).max())
        out[f"f{num_frames}"]["matrix_same_input"] = m2
        assert m2 <= 2e-3, m2
        # DTW: identical path on identical input
        n_sot = len(tk.sot_sequence)
        ti, tj = T.dtw(m, mat[n_sot: n - 1])
        ri, rj = OT.dtw(-mat[n_sot: n - 1].cpu().numpy())
        assert np.array_equal(ti, ri) and np.array_equal(tj, rj), "DTW path differs"
        assert ti[0] == 0 and tj[0] == 0 and ti[-1] == len(text_tokens) and tj[-1] == num_frames // 2 - 1
        # words: same boundaries as the oracle's bookkeeping run on the GPU path
        words = T.find_alignment(m, tk, text_tokens, _bf16(xa).cuda(), num_frames)
        assert len(words) >= 1 and all(0.0 <= wd.start <= wd.end <= num_frames / 100 + 1e-6 for wd in words)
        assert [t for wd in words for t in wd.tokens] == text_tokens  # the trailing EOT "word" is dropped by the zip
        assert all(0.0 <= wd.probability <= 1.0 for wd in words)
        out[f"f{num_frames}"]["n_words"] = len(words)
    # ---- end to end: every segment carries its words, monotonic and inside the file
    mm = _product("micro")
    audio = synth.long_audio(50.0, 3)
    for hst in (None, 2.0):
        r = transcribe(audio, model=mm, temperature=0.0, language="en", sample_len=24, word_timestamps=True,
                       condition_on_previous_text=False, hallucination_silence_threshold=hst,
                       logprob_threshold=None, compression_ratio_threshold=None, no_speech_threshold=None)
        # (with random weights every word is improbable: the hallucination filter may drop all segments)
        assert (len(r["segments"]) >= 1 or hst is not None) and all("words" in s for s in r["segments"])
        n_words = 0
        for s in r["segments"]:
            last = -1.0
            for wd in s["words"]:
                assert wd["start"] <= wd["end"] and wd["end"] <= 51.0 and wd["start"] >= last - 1e-6, (s["id"], wd)
                last = wd["start"]
                n_words += 1
        # words are handed out by token counts, so one may straddle a segment boundary: compare per window
        for seek in sorted({s["seek"] for s in r["segments"]}):
            if seek > 5000 - 4:
                continue  # a tail window of a single attention frame carries no words (timing.py guard)
            segs = [s for s in r["segments"] if s["seek"] == seek]
            joined_words = "".join(wd["word"] for s in segs for wd in s["words"])
            joined_text = "".join(s["text"] for s in segs)
            assert joined_words.replace(" ", "") == joined_text.replace(" ", ""), (seek, joined_words, joined_text)
        out[f"transcribe_hst{hst}"] = (len(r["segments"]), n_words)
        assert n_words >= 1 or hst is not None
    return out


CASES = {
    "logmel_noise": case_logmel_noise,
    "logmel_kinds": case_logmel_kinds,
    "logmel_shapes": case_logmel_shapes,
    "gemm": case_gemm,
    "gemm_vocab": case_gemm_vocab,
    "conv": case_conv,
    "layernorm_embed": case_layernorm_embed,
    "encoder_attention": case_encoder_attention,
    "decoder_attention": case_decoder_attention,
    "splitk_decode_ops": case_splitk_decode_ops,
    "absorbed_cross_attention": case_absorbed_cross_attention,
    "logmel_pcm16": case_logmel_pcm16,
    "ring_attention_edges": case_ring_attention_edges,
    "filter_argmax": case_filter_argmax,
    "sampling_distribution": case_sampling_distribution,
    "encoder_tiny": case_encoder_tiny,
    "decoder_tiny": case_decoder_tiny,
    "decode_tiny": case_decode_tiny,
    "large_v3_parity": case_large_v3_parity,
    "small_batch_step": case_small_batch_step,
    "decode_dual_stream": case_decode_dual_stream,
    "transcribe_micro": case_transcribe_micro,
    "word_alignment": case_word_alignment,
    "config2_logmel_batch1024": case_config2_logmel_batch1024,
    "config3_small_batch64": case_config3_small_batch64,
    "config4_large_v3_batch120": case_config4_large_v3_batch120,
    "config5_turbo_batch256": case_config5_turbo_batch256,
}
