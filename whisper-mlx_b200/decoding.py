"""Per-window decoding: host mirror of `mlx_whisper/decoding.py` (UPSTREAM; reached from
/root/reference/run:3-6; restated in SURVEY.md A.4).

`DecodingOptions`, `DecodingResult`, `decode`, `detect_language` keep the reference names, fields and
error behaviour (beam search raises NotImplementedError there too).  What differs is where the work
runs: the whole step -- decoder forward over a paged KV cache, logit suppression, timestamp rules,
log-softmax, greedy / sampled token choice, sum_logprobs, EOT bookkeeping -- stays on the device
(csrc: b200w_decoder_step), steps are replayed from a CUDA graph, and the host only polls a
`finished` flag every few steps.  The reference syncs to the host every step.
"""
from __future__ import annotations

import ctypes as C
import zlib
from dataclasses import dataclass, field, replace
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .audio import CHUNK_LENGTH, N_FRAMES
from .tokenizer import Tokenizer, get_tokenizer


# kernels launched through CUDA-graph replays (these bypass the library's own launch counter);
# total kernels = b200w_launch_count() + GRAPH_KERNEL_LAUNCHES
GRAPH_KERNEL_LAUNCHES = 0


def total_kernel_launches() -> int:
    return int(_lib.load().b200w_launch_count()) + GRAPH_KERNEL_LAUNCHES


def compression_ratio(text) -> float:
    text_bytes = text.encode("utf-8")
    return len(text_bytes) / len(zlib.compress(text_bytes))


@dataclass(frozen=True)
class DecodingOptions:
    task: str = "transcribe"  # "transcribe" or "translate"
    language: Optional[str] = None  # language that the audio is in; uses detected language if None
    temperature: float = 0.0
    sample_len: Optional[int] = None  # maximum number of tokens to sample
    best_of: Optional[int] = None  # number of independent sample trajectories, if t > 0
    beam_size: Optional[int] = None  # number of beams in beam search, if t == 0
    patience: Optional[float] = None
    length_penalty: Optional[float] = None
    prompt: Optional[Union[str, List[int]]] = None  # for the previous context
    prefix: Optional[Union[str, List[int]]] = None  # to prefix the current context
    suppress_tokens: Optional[Union[str, Iterable[int]]] = "-1"
    suppress_blank: bool = True
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    fp16: bool = True  # 16-bit storage (bf16 on B200), fp32 accumulation
    seed: int = 0  # sampling seed (temperature > 0); not in the reference, which uses the global MLX RNG


@dataclass(frozen=True)
class DecodingResult:
    audio_features: Optional[torch.Tensor]
    language: str
    language_probs: Optional[Dict[str, float]] = None
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    avg_logprob: float = np.nan
    no_speech_prob: float = np.nan
    temperature: float = np.nan
    compression_ratio: float = np.nan


class DecodeSession:
    """Device-side decoding state for a batch of sequences sharing one set of cross K/V slots.

    Owns (through torch) the token buffer, per-sequence counters, the paged self-attention cache with its
    block table, the cross K/V and the logits, and drives `b200w_decoder_step`.  A session is sized by
    (n_audio, n_group, max_tokens); `Whisper.decode_session` keeps sessions alive between batches of the same
    shape so that the buffers (30+ GB of cross K/V for 120 large-v3 windows) and the captured CUDA graphs are
    reused: `load()` re-fills the cross K/V and resets the counters.
    """

    def __init__(self, model, audio_features: Union[torch.Tensor, int], n_group: int = 1,
                 max_tokens: Optional[int] = None, cross_kv: Optional[torch.Tensor] = None):
        self.model = model
        self.lib = model._lib
        dm = model.dims
        dev = model.device
        if isinstance(audio_features, int):
            self.n_audio = audio_features
            audio_features = None
        else:
            if audio_features.ndim == 2:
                audio_features = audio_features[None]
            self.n_audio = audio_features.shape[0]
        self.n_group = n_group
        self.n_seq = self.n_audio * n_group
        B = self.n_seq
        max_tokens = min(max_tokens or dm.n_text_ctx, dm.n_text_ctx)
        ps = model.PAGE_SIZE
        self.max_pages = (max_tokens + ps - 1) // ps
        self.tokens_ld = dm.n_text_ctx + 8
        self.tokens = torch.zeros((B, self.tokens_ld), dtype=torch.int32, device=dev)
        self.n_tokens = torch.zeros(B, dtype=torch.int32, device=dev)
        self.pos = torch.zeros(B, dtype=torch.int32, device=dev)
        self.sum_logprob = torch.zeros(B, dtype=torch.float32, device=dev)
        self.finished = torch.zeros(B, dtype=torch.int32, device=dev)
        self.no_speech = torch.zeros(B, dtype=torch.float32, device=dev)
        d = dm.n_text_state
        n_pages = B * self.max_pages
        self.k_pages = torch.empty((dm.n_text_layer, n_pages, ps, d), dtype=torch.bfloat16, device=dev)
        self.v_pages = torch.empty_like(self.k_pages)
        # page allocator: sequence b owns pages [b * max_pages, (b + 1) * max_pages)
        self.block_table = torch.arange(n_pages, dtype=torch.int32, device=dev).view(B, self.max_pages).contiguous()
        if cross_kv is not None:
            self.cross = cross_kv
        else:
            self.cross = torch.empty((dm.n_text_layer, self.n_audio, dm.n_audio_ctx, 2 * d), dtype=torch.bfloat16, device=dev)
            if audio_features is not None:
                model.cross_kv(audio_features.to(torch.bfloat16), out=self.cross)
        self.cross_slot = (torch.arange(B, dtype=torch.int32, device=dev) // n_group).contiguous()
        ld = model.logits_ld
        self.logits = torch.empty((B, ld), dtype=torch.float32, device=dev)
        self.logits_aux = torch.empty((B, ld), dtype=torch.float32, device=dev)
        self.suppress_bits = torch.zeros((dm.n_vocab + 31) // 32, dtype=torch.int32, device=dev)
        self._ws: Dict[Union[int, str], torch.Tensor] = {}
        self._graphs: Dict[bytes, Tuple[torch.cuda.CUDAGraph, int]] = {}  # keyed by the filter parameters baked in
        self._fp = _lib.FilterParams()
        self.state = _lib.DecodeState(
            B, _lib.ptr(self.tokens), self.tokens_ld, _lib.ptr(self.n_tokens), _lib.ptr(self.pos),
            _lib.ptr(self.sum_logprob), _lib.ptr(self.finished), _lib.ptr(self.no_speech), _lib.ptr(self.k_pages),
            _lib.ptr(self.v_pages), self.k_pages.stride(0), _lib.ptr(self.block_table), self.max_pages, ps,
            _lib.ptr(self.cross), self.cross.stride(0), _lib.ptr(self.cross_slot), _lib.ptr(self.logits),
            _lib.ptr(self.logits_aux), ld, _lib.ptr(self.suppress_bits), None, 0)
        self.xa: Optional[torch.Tensor] = None
        if audio_features is not None and cross_kv is None:
            self._keep_xa(audio_features)

    # -- setup -------------------------------------------------------------------------------------
    def load(self, audio_features: torch.Tensor) -> None:
        """Start a new batch in this session: project the encoder states into the (persistent) cross K/V."""
        if audio_features.ndim == 2:
            audio_features = audio_features[None]
        assert audio_features.shape[0] == self.n_audio
        self.model.cross_kv(audio_features.to(torch.bfloat16), out=self.cross)
        self._keep_xa(audio_features)

    def _keep_xa(self, audio_features: torch.Tensor) -> None:
        """The encoder states behind the cross K/V stay with the session: the absorbed cross-attention (K14) of large
        batches reads them instead of the per-layer K / V."""
        xa = audio_features.to(torch.bfloat16)
        if self.xa is None or self.xa.shape != xa.shape:
            self.xa = torch.empty_like(xa, memory_format=torch.contiguous_format)
        self.xa.copy_(xa)
        self.state.xa = self.xa.data_ptr()
        self.state.xa_slots = self.n_audio

    def set_tokens(self, tokens: torch.Tensor) -> None:
        """Load token histories (B, n); nothing is cached yet."""
        B, n = tokens.shape
        assert B == self.n_seq and n <= self.tokens_ld
        self.tokens[:, :n] = tokens.to(self.tokens.device, torch.int32)
        self.n_tokens.fill_(n)
        self.pos.zero_()
        self.sum_logprob.zero_()
        self.finished.zero_()

    def set_filter(self, fp: "_lib.FilterParams", suppress: Sequence[int]) -> None:
        bits = np.zeros(self.suppress_bits.numel(), dtype=np.uint32)
        for t in suppress:
            bits[t >> 5] |= np.uint32(1 << (t & 31))
        self.suppress_bits.copy_(torch.from_numpy(bits.view(np.int32)))
        self._fp = fp

    def _workspace(self, n_q: int) -> torch.Tensor:
        if n_q not in self._ws:
            nbytes = self.lib.b200w_decoder_workspace_bytes(self.model._handle, self.n_seq, n_q)
            self._ws[n_q] = torch.empty(nbytes, dtype=torch.uint8, device=self.model.device)
        return self._ws[n_q]

    # -- steps -------------------------------------------------------------------------------------
    def _step(self, n_q: int, sot_index: int, select: bool) -> None:
        ws = self._workspace(n_q)
        _lib.check(self.lib.b200w_decoder_step(self.model._handle, C.byref(self.state), n_q, sot_index, int(select),
                                               C.byref(self._fp), _lib.ptr(ws), ws.numel(), _lib.stream()))

    def forward(self, n_q: int) -> torch.Tensor:
        """Teacher forcing: run the decoder over the next n_q already-loaded tokens, return last-token logits."""
        with torch.cuda.device(self.model.device):
            self._step(n_q, 0 if n_q > 1 else -1, False)
        self.pos += n_q
        return self.logits

    def aux_logits(self) -> torch.Tensor:
        return self.logits_aux

    def forward_full(self, n_q: int, probs_first_layer: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Teacher forcing over the n_q loaded tokens with everything word-level alignment needs: the logits of every
        position, (n_seq * n_q, n_vocab) f32, and the cross-attention probabilities of the layers >= probs_first_layer,
        (layers, n_seq, n_q, n_head, 1500) f32 (the decoder half of UPSTREAM Whisper.forward_with_cross_qk)."""
        dm = self.model.dims
        dev = self.model.device
        n_layers = dm.n_text_layer - probs_first_layer
        logits = torch.empty((self.n_seq * n_q, self.model.logits_ld), dtype=torch.float32, device=dev)
        probs = torch.empty((n_layers, self.n_seq, n_q, dm.n_text_head, dm.n_audio_ctx), dtype=torch.float32, device=dev)
        if "full" not in self._ws:  # one buffer for every sequence length (short sequences also carve split-K slabs)
            nbytes = max(self.lib.b200w_decoder_workspace_bytes(self.model._handle, self.n_seq, q) for q in (1, dm.n_text_ctx))
            self._ws["full"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ws = self._ws["full"]
        with torch.cuda.device(dev):
            _lib.check(self.lib.b200w_decoder_forward_full(self.model._handle, C.byref(self.state), n_q, _lib.ptr(ws), ws.numel(),
                                                           _lib.ptr(logits), _lib.ptr(probs), probs_first_layer, _lib.stream()))
        return logits[:, : dm.n_vocab], probs

    def prompt_step(self, n_q: int, sot_index: int) -> None:
        with torch.cuda.device(self.model.device):
            self._step(n_q, sot_index, True)

    def sample_step(self) -> None:
        """One single-token step + token selection, replayed from a CUDA graph after the first capture."""
        global GRAPH_KERNEL_LAUNCHES
        with torch.cuda.device(self.model.device):
            key = bytes(self._fp)
            if key not in self._graphs:
                self._workspace(1)
                before = self.lib.b200w_launch_count()
                g = torch.cuda.CUDAGraph()
                # raw capture: torch.cuda.graph() would synchronise and empty the allocator cache
                s = torch.cuda.Stream(device=self.model.device)
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    g.capture_begin()
                    try:
                        self._step(1, -1, True)
                    finally:
                        g.capture_end()
                torch.cuda.current_stream().wait_stream(s)
                n = self.lib.b200w_launch_count() - before
                GRAPH_KERNEL_LAUNCHES -= n  # the capture itself launched nothing
                self._graphs[key] = (g, n)
            g, n = self._graphs[key]
            self._graph_kernels = n
            g.replay()
            GRAPH_KERNEL_LAUNCHES += n


def _vocab_dir(model) -> Optional[str]:
    return getattr(model, "model_path", None)


class DecodingTask:
    def __init__(self, model, options: DecodingOptions, tokenizer: Optional[Tokenizer] = None):
        """`tokenizer`: the caller's tokenizer (transcribe() passes its own so that segment text, DecodingResult.text
        and the compression ratio all come from one vocabulary); default: built for the model's directory."""
        self.model = model
        language = options.language or "en"
        if tokenizer is None:
            tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language=language,
                                      task=options.task, vocab_dir=_vocab_dir(model))
        self.tokenizer: Tokenizer = tokenizer
        self.options: DecodingOptions = self._verify_options(options)

        self.n_group: int = options.beam_size or options.best_of or 1
        self.n_ctx: int = model.dims.n_text_ctx
        self.sample_len: int = options.sample_len or model.dims.n_text_ctx // 2

        self.sot_sequence: Tuple[int] = tokenizer.sot_sequence
        if self.options.without_timestamps:
            self.sot_sequence = tokenizer.sot_sequence_including_notimestamps

        self.initial_tokens: Tuple[int] = self._get_initial_tokens()
        self.sample_begin: int = len(self.initial_tokens)
        self.sot_index: int = self.initial_tokens.index(tokenizer.sot)
        if options.beam_size is not None:
            raise NotImplementedError("Beam search decoder is not yet implemented")

    def _verify_options(self, options: DecodingOptions) -> DecodingOptions:
        if options.beam_size is not None and options.best_of is not None:
            raise ValueError("beam_size and best_of can't be given together")
        if options.temperature == 0:
            if options.best_of is not None:
                raise ValueError("best_of with greedy sampling (T=0) is not compatible")
        if options.patience is not None and options.beam_size is None:
            raise ValueError("patience requires beam_size to be given")
        if options.length_penalty is not None and not (0 <= options.length_penalty <= 1):
            raise ValueError("length_penalty (alpha) should be a value between 0 and 1")
        return options

    def _get_initial_tokens(self) -> Tuple[int]:
        tokens = list(self.sot_sequence)
        if prefix := self.options.prefix:
            prefix_tokens = self.tokenizer.encode(" " + prefix.strip()) if isinstance(prefix, str) else prefix
            if self.sample_len is not None:
                max_prefix_len = self.n_ctx // 2 - self.sample_len
                prefix_tokens = prefix_tokens[-max_prefix_len:]
            tokens = tokens + list(prefix_tokens)
        if prompt := self.options.prompt:
            prompt_tokens = self.tokenizer.encode(" " + prompt.strip()) if isinstance(prompt, str) else prompt
            tokens = [self.tokenizer.sot_prev] + list(prompt_tokens[-(self.n_ctx // 2 - 1):]) + tokens
        return tuple(tokens)

    def _get_suppress_tokens(self) -> Tuple[int]:
        suppress_tokens = self.options.suppress_tokens
        if isinstance(suppress_tokens, str):
            suppress_tokens = [int(t) for t in suppress_tokens.split(",")]
        if -1 in suppress_tokens:
            suppress_tokens = [t for t in suppress_tokens if t >= 0]
            suppress_tokens.extend(self.tokenizer.non_speech_tokens)
        elif suppress_tokens is None or len(suppress_tokens) == 0:
            suppress_tokens = []
        else:
            assert isinstance(suppress_tokens, list), "suppress_tokens must be a list"
        tk = self.tokenizer
        suppress_tokens.extend([tk.transcribe, tk.translate, tk.sot, tk.sot_prev, tk.sot_lm])
        if tk.no_speech is not None:
            suppress_tokens.append(tk.no_speech)
        return tuple(sorted(set(suppress_tokens)))

    def _filter_params(self, sess: DecodeSession) -> "_lib.FilterParams":
        tk, dm, opt = self.tokenizer, self.model.dims, self.options
        max_init = -1
        if opt.max_initial_timestamp is not None:
            precision = CHUNK_LENGTH / dm.n_audio_ctx  # usually 0.02 seconds
            max_init = round(opt.max_initial_timestamp / precision)
        return _lib.FilterParams(
            n_vocab=dm.n_vocab, logits_ld=self.model.logits_ld, sample_begin=self.sample_begin, eot=tk.eot,
            blank=tk.encode(" ")[0], no_timestamps=tk.no_timestamps, timestamp_begin=tk.timestamp_begin,
            no_speech=tk.no_speech, max_initial_timestamp_index=max_init,
            apply_timestamp_rules=0 if opt.without_timestamps else 1, suppress_blank=1 if opt.suppress_blank else 0,
            tokens_ld=sess.tokens_ld, temperature=float(opt.temperature), seed=int(opt.seed))

    # ---------------------------------------------------------------------------------------------
    def run_features(self, audio_features: torch.Tensor, cross_kv: Optional[torch.Tensor] = None,
                     poll_every: int = 8) -> List[DecodingResult]:
        """Decode a batch of encoder states (n_audio, 1500, d)."""
        tk = self.tokenizer
        n_audio = audio_features.shape[0]
        n0 = len(self.initial_tokens)
        max_tokens = min(self.n_ctx, n0 + self.sample_len)
        model = self.model
        # Two half-batches decoded on two streams: while one half streams its cross K/V (HBM-bound), the other
        # half's small GEMM / LayerNorm / self-attention kernels (latency-bound) run beside it.
        n_streams = 2 if (cross_kv is None and model.decode_streams >= 2 and n_audio >= 16) else 1
        bounds = [(n_audio * i) // n_streams for i in range(n_streams + 1)]
        sessions: List[DecodeSession] = []
        for i in range(n_streams):
            part = audio_features[bounds[i]: bounds[i + 1]]
            if cross_kv is not None:
                sess = DecodeSession(model, part, self.n_group, max_tokens=max_tokens, cross_kv=cross_kv)
            else:
                sess = model.decode_session(part.shape[0], self.n_group, max_tokens, slot=i)
                sess.load(part)
            init = torch.tensor(self.initial_tokens, dtype=torch.int32).repeat(sess.n_seq, 1)
            sess.set_tokens(init)
            sess.set_filter(self._filter_params(sess), self._get_suppress_tokens())
            sessions.append(sess)

        for sess in sessions:
            sess.prompt_step(n0, self.sot_index)
        steps = 1
        max_steps = min(self.sample_len, self.n_ctx + 1 - n0)
        if n_streams == 1:
            sess = sessions[0]
            while steps < max_steps:
                sess.sample_step()
                steps += 1
                if steps % poll_every == 0 and bool(sess.finished.all().item()):
                    break
        else:
            cur = torch.cuda.current_stream(model.device)
            streams = model.side_streams(n_streams)
            for st in streams:
                st.wait_stream(cur)
            while steps < max_steps:
                for st, sess in zip(streams, sessions):
                    with torch.cuda.stream(st):
                        sess.sample_step()
                steps += 1
                if steps % poll_every == 0:
                    for st in streams:
                        st.synchronize()
                    if all(bool(sess.finished.all().item()) for sess in sessions):
                        break
            for st in streams:
                cur.wait_stream(st)

        n_tok = int(sessions[0].n_tokens[0].item())
        tokens = torch.cat([sess.tokens[:, :n_tok] for sess in sessions], 0).cpu().numpy()
        sum_logprobs = torch.cat([sess.sum_logprob for sess in sessions], 0).cpu().numpy().astype(np.float64)
        no_speech = torch.cat([sess.no_speech for sess in sessions], 0).cpu().numpy()[:: self.n_group]

        tokens = tokens.reshape(n_audio, self.n_group, n_tok)
        sum_logprobs = sum_logprobs.reshape(n_audio, self.n_group)
        results = []
        for a in range(n_audio):
            cands = []
            for t in tokens[a]:
                seq = t[self.sample_begin:].tolist() + [tk.eot]  # finalize() pads one EOT
                cands.append(seq[: seq.index(tk.eot)])
            # MaximumLikelihoodRanker: sum_logprob / length (or the Google NMT penalty)
            def score(i):
                length = len(cands[i])
                lp = self.options.length_penalty
                penalty = length if lp is None else ((5 + length) / 6) ** lp
                return sum_logprobs[a, i] / penalty if penalty > 0 else -np.inf

            best = int(np.argmax([score(i) for i in range(self.n_group)])) if self.n_group > 1 else 0
            toks = cands[best]
            text = tk.decode(toks).strip()
            results.append(DecodingResult(
                audio_features=audio_features[a], language=self.options.language or "en", tokens=toks, text=text,
                avg_logprob=float(sum_logprobs[a, best]) / (len(toks) + 1), no_speech_prob=float(no_speech[a]),
                temperature=self.options.temperature, compression_ratio=compression_ratio(text)))
        return results

    def run(self, mel: torch.Tensor) -> List[DecodingResult]:
        model = self.model
        if mel.ndim == 2:
            mel = mel[None]
        if mel.shape[-2:] == (model.dims.n_audio_ctx, model.dims.n_audio_state):
            audio_features = mel  # encoded audio features are given; skip audio encoding
        else:
            audio_features = model.embed_audio(mel)
        return self.run_features(audio_features)


def decode(model, mel: torch.Tensor, options: DecodingOptions = DecodingOptions(), **kwargs):
    """Decode 30-second window(s): mel (3000, n_mels) or (B, 3000, n_mels) -> DecodingResult or list."""
    mel = torch.as_tensor(mel)
    if single := mel.ndim == 2:
        mel = mel[None]
    if kwargs:
        options = replace(options, **kwargs)
    if options.language is None:
        lang_tokens, lang_probs = detect_language(model, mel)
        tk = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, vocab_dir=_vocab_dir(model))
        codes = [tk.language_code(t) for t in lang_tokens]
        # the reference decodes each window with its own language; batches here share one prompt
        if len(set(codes)) > 1:
            out = []
            for i, c in enumerate(codes):
                out.extend(DecodingTask(model, replace(options, language=c)).run(mel[i: i + 1]))
            return out[0] if single else out
        options = replace(options, language=codes[0])
    result = DecodingTask(model, options).run(mel)
    return result[0] if single else result


def detect_language(model, mel: torch.Tensor, tokenizer: Optional[Tokenizer] = None):
    """Detect the spoken language: one decoder step on [sot], argmax / softmax over the language tokens.

    Returns (language_tokens list[int], language_probs list[dict]); scalars for a single window.
    """
    if tokenizer is None:
        tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, vocab_dir=_vocab_dir(model))
    if tokenizer.language is None or tokenizer.language_token not in tokenizer.sot_sequence:
        raise ValueError("This model doesn't have language tokens so it can't perform lang id")
    mel = torch.as_tensor(mel)
    single = mel.ndim == 2
    if single:
        mel = mel[None]
    if mel.shape[-2:] != (model.dims.n_audio_ctx, model.dims.n_audio_state):
        mel = model.embed_audio(mel)
    n_audio = mel.shape[0]
    sess = DecodeSession(model, mel, 1, max_tokens=16)
    sess.set_tokens(torch.full((n_audio, 1), tokenizer.sot, dtype=torch.int32))
    logits = sess.forward(1)
    n_lang = len(tokenizer.all_language_tokens)
    lang_begin = min(tokenizer.all_language_tokens)
    tok = torch.empty(n_audio, dtype=torch.int32, device=model.device)
    probs = torch.empty((n_audio, n_lang), dtype=torch.float32, device=model.device)
    with torch.cuda.device(model.device):
        _lib.check(model._lib.b200w_detect_language(_lib.ptr(logits), model.logits_ld, n_audio, lang_begin, n_lang,
                                                    _lib.ptr(tok), _lib.ptr(probs), _lib.stream()))
    tok_h, probs_h = tok.cpu().tolist(), probs.cpu().numpy()
    codes = tokenizer.all_language_codes
    language_probs = [{c: float(probs_h[i, j]) for j, c in enumerate(codes)} for i in range(n_audio)]
    if single:
        return tok_h[0], language_probs[0]
    return tok_h, language_probs
