"""Oracle: per-window decoding (test infrastructure only; see oracle/__init__.py).

Restates `mlx_whisper/decoding.py` (UPSTREAM, not under /root/reference; call site
/root/reference/run:3-6) per SURVEY.md Appendix A.4: logit filters (SuppressBlank,
SuppressTokens, ApplyTimestampRules), GreedyDecoder, DecodingTask main loop, detect_language.
The timestamp rules are cross-pinned against transformers/generation/logits_process.py:1996-2041
in tests/test_oracle_vs_hf.py.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import model as M
from .tokens import TokenIds, decode_text, LANGUAGE_CODES

NEG_INF = float("-inf")


def compression_ratio(text: str) -> float:
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


@dataclass
class DecodingResult:
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    language: str = "en"
    avg_logprob: float = float("nan")
    no_speech_prob: float = float("nan")
    temperature: float = 0.0
    compression_ratio: float = float("nan")
    sum_logprob: float = float("nan")
    # per sampled token: gap between the two best filtered logits (how close the greedy choice was to flipping)
    margins: List[float] = field(default_factory=list)


# ------------------------------------------------------------------------------ logit filters

def suppress_blank(logits: np.ndarray, n_tokens: int, sample_begin: int, ids: TokenIds):
    """At the first sampled position forbid " " (220) and EOT."""
    if n_tokens == sample_begin:
        logits[:, [ids.blank, ids.eot]] = NEG_INF


def suppress_tokens(logits: np.ndarray, suppress: Sequence[int]):
    logits[:, list(suppress)] = NEG_INF


def apply_timestamp_rules(logits: np.ndarray, tokens: np.ndarray, sample_begin: int, ids: TokenIds,
                          max_initial_timestamp_index: Optional[int] = 50):
    """Timestamp grammar. `logits` (B, V) f32 is edited in place; `tokens` (B, n)."""
    tb = ids.timestamp_begin
    logits[:, ids.no_timestamps] = NEG_INF
    for k in range(tokens.shape[0]):
        seq = tokens[k, sample_begin:].tolist()
        last_ts = len(seq) >= 1 and seq[-1] >= tb
        penult_ts = len(seq) < 2 or seq[-2] >= tb
        if last_ts:
            if penult_ts:
                logits[k, tb:] = NEG_INF  # a closed pair must be followed by text (or EOT)
            else:
                logits[k, : ids.eot] = NEG_INF  # an opening timestamp after text: no plain text next
        stamps = [t for t in seq if t >= tb]
        if stamps:
            floor = stamps[-1] if (last_ts and not penult_ts) else stamps[-1] + 1
            logits[k, tb:floor] = NEG_INF  # timestamps never decrease; segments have nonzero length
    if tokens.shape[1] == sample_begin:
        logits[:, :tb] = NEG_INF
        if max_initial_timestamp_index is not None:
            logits[:, tb + max_initial_timestamp_index + 1:] = NEG_INF
    lp = torch.log_softmax(torch.from_numpy(logits).float(), dim=-1)
    for k in range(tokens.shape[0]):
        ts_lp = torch.logsumexp(lp[k, tb:], dim=-1)
        text_max = lp[k, :tb].max()
        if ts_lp > text_max:
            logits[k, :tb] = NEG_INF


def filter_logits(logits: np.ndarray, tokens: np.ndarray, sample_begin: int, ids: TokenIds,
                  suppress: Sequence[int], without_timestamps: bool = False,
                  max_initial_timestamp_index: Optional[int] = 50, blank: bool = True):
    """SuppressBlank -> SuppressTokens -> ApplyTimestampRules, in upstream order."""
    if blank:
        suppress_blank(logits, tokens.shape[1], sample_begin, ids)
    suppress_tokens(logits, suppress)
    if not without_timestamps:
        apply_timestamp_rules(logits, tokens, sample_begin, ids, max_initial_timestamp_index)
    return logits


def greedy_update(tokens: np.ndarray, logits: np.ndarray, sum_logprobs: np.ndarray, eot: int,
                  temperature: float = 0.0, generator: Optional[torch.Generator] = None):
    """GreedyDecoder.update."""
    lt = torch.from_numpy(logits).float()
    if temperature == 0.0:
        nxt = lt.argmax(dim=-1)
    else:
        nxt = torch.multinomial(torch.softmax(lt / temperature, dim=-1), 1, generator=generator)[:, 0]
    lp = torch.log_softmax(lt, dim=-1)
    cur = lp[torch.arange(lp.shape[0]), nxt].numpy()
    alive = tokens[:, -1] != eot
    sum_logprobs += cur * alive
    nxt = nxt.numpy().astype(tokens.dtype)
    nxt[~alive] = eot
    tokens = np.concatenate([tokens, nxt[:, None]], axis=1)
    return tokens, bool((tokens[:, -1] == eot).all())


# ------------------------------------------------------------------------------ language id

@torch.no_grad()
def detect_language(w, dims: M.ModelDimensions, xa: torch.Tensor, policy="fp32"):
    """One decoder step on [sot]; argmax over the language tokens."""
    ids = TokenIds(dims.n_vocab)
    B = xa.shape[0]
    logits, _ = M.decoder_forward(w, dims, torch.full((B, 1), ids.sot, dtype=torch.long), xa, policy=policy)
    logits = logits[:, 0].clone()
    mask = torch.ones(dims.n_vocab, dtype=torch.bool)
    mask[ids.language_begin: ids.language_begin + ids.num_languages] = False
    logits[:, mask] = NEG_INF
    lang_tokens = logits.argmax(dim=-1)
    probs = torch.softmax(logits, dim=-1)
    out = []
    for b in range(B):
        out.append({LANGUAGE_CODES[j]: probs[b, ids.language_begin + j].item() for j in range(ids.num_languages)})
    return lang_tokens.tolist(), out


# ------------------------------------------------------------------------------ decoding task

@torch.no_grad()
def decode(w, dims: M.ModelDimensions, mel: torch.Tensor, *, language: str = "en", task: str = "transcribe",
           temperature: float = 0.0, sample_len: Optional[int] = None, prompt: Sequence[int] = (),
           without_timestamps: bool = False, max_initial_timestamp: Optional[float] = 1.0,
           suppress_blank_: bool = True, policy: str = "fp32", audio_features: Optional[torch.Tensor] = None,
           seed: int = 0, return_logits: bool = False):
    """DecodingTask.run for a batch of windows. mel (B, 3000, n_mels) -> list[DecodingResult]."""
    ids = TokenIds(dims.n_vocab)
    n_ctx = dims.n_text_ctx
    sample_len = sample_len or n_ctx // 2
    sot_seq = list(ids.sot_sequence(language, task))
    if without_timestamps:
        sot_seq.append(ids.no_timestamps)
    initial = []
    if len(prompt) > 0:
        initial = [ids.sot_prev] + list(prompt)[-(n_ctx // 2 - 1):]
    initial = initial + sot_seq
    sample_begin = len(initial)
    sot_index = initial.index(ids.sot)
    suppress = ids.suppress_set()
    max_init_idx = None
    if max_initial_timestamp is not None:
        max_init_idx = round(max_initial_timestamp / (30.0 / dims.n_audio_ctx))

    xa = audio_features if audio_features is not None else M.encoder_forward(w, dims, mel, policy=policy)
    B = xa.shape[0]
    tokens = np.tile(np.array(initial, dtype=np.int64)[None], (B, 1))
    sum_lp = np.zeros(B, dtype=np.float32)
    no_speech = np.full(B, np.nan, dtype=np.float32)
    gen = torch.Generator().manual_seed(seed)
    cache = None
    kept_logits = []
    step_margins = []
    for i in range(sample_len):
        inp = tokens if i == 0 else tokens[:, -1:]
        logits, cache = M.decoder_forward(w, dims, torch.from_numpy(inp), xa, cache, policy=policy)
        if i == 0:
            no_speech = torch.softmax(logits[:, sot_index].float(), dim=-1)[:, ids.no_speech].numpy()
        step = logits[:, -1].float().numpy().copy()
        if return_logits:
            kept_logits.append(step.copy())
        filter_logits(step, tokens, sample_begin, ids, suppress, without_timestamps, max_init_idx, suppress_blank_)
        top2 = torch.from_numpy(step).topk(2, dim=-1).values
        step_margins.append((top2[:, 0] - top2[:, 1]).tolist())
        tokens, done = greedy_update(tokens, step, sum_lp, ids.eot, temperature, gen)
        if done or tokens.shape[-1] > n_ctx:
            break
    tokens = np.concatenate([tokens, np.full((B, 1), ids.eot, dtype=tokens.dtype)], axis=1)  # finalize
    results = []
    for b in range(B):
        t = tokens[b, sample_begin:].tolist()
        t = t[: t.index(ids.eot)]
        text = decode_text(t, ids.timestamp_begin).strip()
        results.append(DecodingResult(
            tokens=t, text=text, language=language, sum_logprob=float(sum_lp[b]),
            avg_logprob=float(sum_lp[b]) / (len(t) + 1), no_speech_prob=float(no_speech[b]),
            temperature=temperature, compression_ratio=compression_ratio(text),
            margins=[float(sm[b]) for sm in step_margins[: len(t) + 1]]))
    if return_logits:
        return results, kept_logits
    return results
