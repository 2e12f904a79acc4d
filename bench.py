#!/usr/bin/env python
"""Benchmark of the Whisper transcription hot path on B200 (contract: see DESIGN.md, "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak] [--mode batched|exact]
                    [--config 4|2|3|5] [--model large-v3] [--hours 1.0]

Default workload = BASELINE.json configs[3] ("config 4"): whisper-large-v3 (128 mel, 32+32 layers, random-init, bf16)
greedy transcription of ONE hour of synthetic 16 kHz audio = 120 fixed 30 s windows, through the public
`transcribe(audio, ...)` API in fixed-window batched mode.  One step = one pass over the whole hour.
  value : RTFx (audio seconds per second of device time) with the audio already resident in HBM.
  e2e   : the same call with the audio in pinned HOST memory: the H2D copy of the samples and the D2H read of the
          decoded tokens are inside the timed region.
Multi-GPU (one process per GPU, torchrun): `--scaling strong` (default, what BASELINE configs[3] states) shards the
hour's windows over the ranks through the product's own `transcribe(rank=, world_size=)` path -- every rank computes
the file-wide log-mel, decodes its block of windows and the per-window segments are gathered on the host
(all_gather_object; no collective on the data path); RTFx = 3600 / max-over-ranks time.  The weak-scaling number
(every rank its own hour, RTFx = N * 3600 / max time) is measured in the same run and reported under "weak".
`--mode exact` times the reference's sequential seek loop (batch 1, the default of transcribe() and the CLI) on
`--exact-seconds` of audio; the default run reports it under "exact_mode".
`--config 2|3|5` print the line of the other BASELINE configurations (log-mel batch 1024; whisper-small batch 64 and
large-v3-turbo batch 256 decode steps against their HBM floors).
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, the only other place that may
execute it) on the host cores: the K timed steps together are ONE full 30 s window (log-mel + encoder + all 224
greedy decode steps, split evenly over the steps), so ms_per_step * steps is what really ran.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# random-init weights and no vocabulary file in this image: token ids are what is measured and compared
os.environ.setdefault("B200W_ALLOW_SURROGATE", "1")

METRIC = "large-v3 RTFx (audio-s/s)"
UNIT = "audio-s/s"
SAMPLE_LEN = 224


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def workload_config(model: str, hours: float, window_batch: int, world: int = 1, scaling: str = "strong") -> dict:
    """The `config` object shared by the product arm and the reference arm (same workload string)."""
    n_windows = int(np.ceil(hours * 3600 / 30))
    return {"workload": f"BASELINE configs[3]: whisper-{model} (random-init) greedy transcription of {hours:g} h synthetic 16 kHz audio "
                        f"= {n_windows} fixed 30 s windows, sample_len {SAMPLE_LEN}, temperature 0, no fallback",
            "windows": n_windows, "sample_len": SAMPLE_LEN}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except FileNotFoundError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ======================================================================================== product arm
def build_model(name: str, seed: int, device: str):
    import torch
    from tools import synth
    from whisper_mlx_b200.whisper import ModelDimensions, Whisper

    dims = synth.DIMS[name]
    weights = dict(synth.random_weights(dims, seed, device=device))
    model = Whisper(ModelDimensions(**dims), weights, device=device)
    return model, weights


def make_audio(hours: float, seed: int) -> np.ndarray:
    from tools import synth

    return synth.long_audio(hours * 3600.0, seed)


def _timed(fn, n):
    import torch

    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


def decoder_param_bytes(dm) -> float:
    """bf16 bytes of everything a single-token decoder step streams besides K/V: the layers' matrices and the tied
    embedding (logits GEMM); cross K/V projection weights are not read in a step."""
    d, L, V = dm.n_text_state, dm.n_text_layer, dm.n_vocab
    per_layer = 3 * d * d + d * d + d * d + d * d + 8 * d * d  # qkv, out, cross q, cross out, mlp1 + mlp2
    return 2.0 * (L * per_layer + V * d)


def decode_step_probe(model, batch: int, peaks, steps: int = 32, position: int = 32):
    """One greedy single-token step (CUDA-graph replay) at `batch` sequences: measured time against its HBM floor
    (decoder weights + every sequence's cross K/V + the self K/V cached so far, each read once), and the per-kernel shares
    of one eager step (CUDA events around every launch of the library)."""
    import torch
    from whisper_mlx_b200._lib import kernel_profile
    from whisper_mlx_b200.decoding import DecodeSession, DecodingOptions, DecodingTask

    dm = model.dims
    xa = torch.randn(batch, dm.n_audio_ctx, dm.n_audio_state, device=model.device).bfloat16()
    task = DecodingTask(model, DecodingOptions(language="en"))
    sess = DecodeSession(model, xa, 1, max_tokens=3 + SAMPLE_LEN)
    sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(batch, 1))
    sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
    sess.prompt_step(len(task.initial_tokens), task.sot_index)
    for _ in range(max(position - steps // 2, 1)):
        sess.sample_step()
    t = _timed(lambda i=0: sess.sample_step(), steps)
    pos_mid = 3 + max(position - steps // 2, 1) + steps // 2
    d, L = dm.n_text_state, dm.n_text_layer
    floor_bytes = decoder_param_bytes(dm) + batch * L * dm.n_audio_ctx * 2 * d * 2.0 + batch * L * pos_mid * 2 * d * 2.0
    with kernel_profile() as prof:
        sess._step(1, -1, True)
    shares = {k: v["total_ms"] for k, v in prof.result.items()}
    tot = sum(shares.values()) or 1.0
    out = {"batch": batch, "step_ms": t * 1e3, "hbm_floor_ms": floor_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3,
           "frac_of_hbm_floor": floor_bytes / (peaks["hbm_gbs"] * 1e9) / t, "tokens_per_s": batch / t,
           "algorithmic_bytes_per_step": floor_bytes, "graph_kernels_per_step": sess._graph_kernels,
           "eager_kernel_share": {k: v / tot for k, v in sorted(shares.items(), key=lambda kv: -kv[1])}}
    del sess, xa
    return out


def kernel_rooflines(model, peaks, n_windows: int, with_logmel: bool = True):
    """Per-kernel achieved vs. roofline at the workload's shapes, timed with CUDA events on the launch stream."""
    import torch
    from whisper_mlx_b200 import _lib as L

    lib = L.load()
    dm = model.dims
    d, T, H = dm.n_text_state, dm.n_audio_ctx, dm.n_text_head
    out = {}

    # K8 cross-attention: one launch per decoder layer per step; each launch streams its layer's K|V once
    Lr = min(dm.n_text_layer, 8)
    ckv = torch.empty((Lr, n_windows, T, 2 * d), dtype=torch.bfloat16, device=model.device).normal_()
    q = torch.randn(n_windows, 1, d, device=model.device).bfloat16()
    o = torch.empty_like(q)
    slot = torch.arange(n_windows, dtype=torch.int32, device=model.device)
    t = _timed(lambda i=0: L.check(lib.b200w_decoder_cross_attention(L.ptr(q), n_windows, 1, H, L.ptr(ckv[i % Lr]), T * 2 * d, T,
                                                                     L.ptr(slot), L.ptr(o), L.stream())), 4 * Lr)
    bytes_alg = n_windows * T * 2 * d * 2  # K and V rows of every window, bf16, read once
    # DRAM traffic per launch from the ncu --set full capture of this kernel at 120 windows
    # (profiles/r02_ncu_cross_ring.json), scaled to this launch's windows
    traffic = CROSS_DRAM_BYTES_AT_120 * n_windows / 120.0 if d == 1280 else None
    # The kernel only READS: besides the copy bandwidth of MEASURED_PEAKS.json (half writes) it is held against a
    # read-only stream of the same number of bytes, timed here the same way (b200w_debug_read_stream).
    import ctypes as C

    rs = lib.b200w_debug_read_stream
    rs.restype, rs.argtypes = C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    sink = torch.zeros(2, dtype=torch.int64, device=model.device)
    t_rd = _timed(lambda i=0: L.check(rs(L.ptr(ckv[i % Lr]), bytes_alg, L.ptr(sink), L.stream())), 4 * Lr)
    read_gbs = bytes_alg / t_rd / 1e9
    out["decoder_cross_attention"] = {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                      "frac": bytes_alg / t / 1e9 / peaks["hbm_gbs"], "traffic": traffic, "launch_ms": t * 1e3,
                                      "algorithmic_bytes_per_launch": bytes_alg, "read_only_stream_gbs": read_gbs,
                                      "frac_of_read_only_stream": t_rd / t,
                                      "note": "peak = copy bandwidth (MEASURED_PEAKS.json); read_only_stream_gbs = the same bytes "
                                              "read linearly by a load-only kernel at the same launch size, timed here"}
    del ckv
    # K5 / K5b encoder GEMMs, all four shapes of a block: M = windows * 1500 rows, L2 flushed by the operand sizes
    M = min(n_windows, 40) * T
    sus = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
    shapes = {"qkv": (3 * d, d, False, False), "out": (d, d, True, False), "mlp1": (4 * d, d, False, True), "mlp2": (d, 4 * d, True, False)}
    tot_fl, tot_t = 0.0, 0.0
    for name, (N, K, resid, gelu) in shapes.items():
        a = torch.randn(M, K, device=model.device).bfloat16()
        w = (torch.randn(N, K, device=model.device) / K ** 0.5).bfloat16()
        bias = torch.zeros(N, device=model.device)
        c = torch.empty((M, N), dtype=torch.float32 if resid else torch.bfloat16, device=model.device)
        flags = (1 if gelu else 0) | (2 if resid else 0)
        t = _timed(lambda i=0: L.check(lib.b200w_gemm_bf16(L.ptr(a), K, L.ptr(w), L.ptr(c), N, L.ptr(bias), L.ptr(c) if resid else None,
                                                           M, N, K, flags, L.stream())), 10)
        fl = 2.0 * M * N * K
        tot_fl += fl
        tot_t += t
        out[f"encoder_gemm_{name}"] = {"bound": "tensor", "achieved": fl / t / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                       "frac": fl / t / 1e12 / peaks["bf16_tflops"], "frac_of_sustained": fl / t / 1e12 / sus,
                                       "traffic": None, "launch_ms": t * 1e3, "algorithmic_flops_per_launch": fl}
        del a, w, c
    # K6 encoder attention at the same number of windows
    Bw = min(n_windows, 40)
    qkv = torch.randn(Bw, T, 3 * d, device=model.device).bfloat16()
    ao = torch.empty((Bw, T, d), dtype=torch.bfloat16, device=model.device)
    t = _timed(lambda i=0: L.check(lib.b200w_encoder_attention(L.ptr(qkv), Bw, T, dm.n_audio_head, L.ptr(ao), L.stream())), 10)
    fl = 4.0 * Bw * dm.n_audio_head * T * T * 64
    tot_fl += fl
    tot_t += t
    out["encoder_attention"] = {"bound": "tensor", "achieved": fl / t / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                "frac": fl / t / 1e12 / peaks["bf16_tflops"], "frac_of_sustained": fl / t / 1e12 / sus, "traffic": None,
                                "launch_ms": t * 1e3, "algorithmic_flops_per_launch": fl}
    out["encoder_block_aggregate"] = {"bound": "tensor", "achieved": tot_fl / tot_t / 1e12, "peak": sus, "unit": "TFLOP/s",
                                      "frac": tot_fl / tot_t / 1e12 / sus, "traffic": None, "launch_ms": tot_t * 1e3,
                                      "note": "FLOP-weighted over the 4 GEMM shapes + attention of one encoder block, vs the SUSTAINED bf16 peak"}
    del qkv, ao
    if with_logmel:
        out.update(logmel_rooflines(peaks, model.device))
    return out


# dram__bytes_read.sum + dram__bytes_write.sum of one decoder_cross_attention_ring_kernel launch at 120 windows (ncu --set full)
CROSS_DRAM_BYTES_AT_120 = 921_962_496 + 4_662_528

# FP32-issue ceiling of the log-mel kernel: warp-instructions per frame from the ncu source view of the shipped kernel
# (profiles/) against 4 warp-instructions per clock per SM on 148 SMs
LOGMEL_WARP_INSTR_PER_FRAME = {80: 419.0, 128: 445.0}


def logmel_rooflines(peaks, device, batch: int = 1024, clocks_mhz: float = None):
    """K1 at BASELINE config 2 (batch 1024 x 30 s), f32 in / f32 out including the clamp: HBM bound and FP32-issue bound."""
    import torch
    from whisper_mlx_b200.audio import log_mel_spectrogram

    x = torch.randn(batch, 480000, device=device) * 0.1
    out = {}
    table = LOGMEL_WARP_INSTR_PER_FRAME
    p = os.path.join(REPO, "profiles", "logmel_instr_per_frame.json")
    if os.path.exists(p):
        table = {int(k): float(v) for k, v in json.load(open(p)).items()}
    for n_mels in (80, 128):
        res = torch.empty((batch, 3000, n_mels), dtype=torch.float32, device=device)  # (the result buffer is reused)
        for _ in range(2):  # (lazy one-time setup of the cooperative launch, tables)
            log_mel_spectrogram(x, n_mels=n_mels, out=res)
        t = _timed(lambda i=0: log_mel_spectrogram(x, n_mels=n_mels, out=res), 10)
        del res
        bytes_alg = batch * (4 * 480000 + 4 * 3000 * n_mels)
        frames_per_s = batch * 3000 / t
        sm_hz = (clocks_mhz or 1900.0) * 1e6
        issue_ceiling = 148 * 4 * sm_hz / table[n_mels]  # frames/s if every issue slot of every SM held a useful instruction
        out[f"logmel_{n_mels}"] = {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": bytes_alg / t / 1e9 / peaks["hbm_gbs"],
                                   # dram read + write of the ncu --set full capture at 1024 windows (profiles/r02_ncu_full_summaries.json)
                                   "traffic": (1_968_447_000 + 1_537_237_000) * batch / 1024.0 if n_mels == 128 else None,
                                   "launch_ms": t * 1e3, "frames_per_s": frames_per_s, "algorithmic_bytes_per_launch": bytes_alg,
                                   "fp32_issue_ceiling_frames_per_s": issue_ceiling, "fp32_issue_frac": frames_per_s / issue_ceiling,
                                   "warp_instr_per_frame": table[n_mels], "issue_clock_mhz": sm_hz / 1e6}
    return out


def _segments_key(res):
    return [(s["seek"], tuple(s["tokens"]), round(s["start"], 6), round(s["end"], 6)) for s in res["segments"]]


def run_product(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (B200); there is no CPU fallback for the product arm")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))

    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.decoding import total_kernel_launches

    peaks = load_peaks()
    model, weights = build_model(args.model, 0, device)
    n_windows = int(np.ceil(args.hours * 3600 / 30))
    kw = dict(model=model, temperature=0.0, condition_on_previous_text=False, language="en",
              window_batch=args.window_batch, encoder_batch=args.encoder_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(audio, steps, **extra):
        res = None
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res = transcribe(audio, **kw, **extra)
        e1.record()
        torch.cuda.synchronize()
        dev_s = e0.elapsed_time(e1) * 1e-3
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([dev_s, wall], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), res

    exact = args.mode == "exact"
    if exact:
        # the reference's sequential seek loop, batch 1 (what `./run` / the CLI default does): replicas only across ranks
        kw.update(window_batch=0)
        seconds = args.exact_seconds
        audio_host = torch.from_numpy(make_audio(seconds / 3600.0, 100 + rank)).pin_memory()
        shard = {}
        total_audio_s = seconds * world
        scaling = "weak"
    else:
        strong = args.scaling == "strong"
        # strong: ONE hour, the same file on every rank, windows sharded; weak: every rank its own hour
        audio_host = torch.from_numpy(make_audio(args.hours, 100 if strong else 100 + rank)).pin_memory()
        shard = dict(rank=rank, world_size=world) if (strong and world > 1) else {}
        total_audio_s = args.hours * 3600.0 * (1 if strong else world)
        scaling = "strong" if strong else "weak"
    audio_dev = audio_host.to(device)

    for _ in range(args.warmup):
        transcribe(audio_dev, **kw, **shard)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = total_kernel_launches()
    dev_s, wall_s, res = run(audio_dev, args.steps, **shard)
    launches = total_kernel_launches() - launches0
    e2e_dev_s, e2e_wall_s, res_h = run(audio_host, args.steps, **shard)
    clocks = sampler.stop()
    n_tokens = sum(len(s["tokens"]) for s in res["segments"])

    extras = {}
    if not exact and world > 1 and shard:
        # the sharded result must be the single-GPU result: every rank decodes the whole hour once, outside the timed region
        full = transcribe(audio_dev, **kw)
        same = _segments_key(full) == _segments_key(res) == _segments_key(res_h)
        flag = torch.tensor([1 if same else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        extras["sharded_equals_single_gpu"] = bool(flag.item())
        if not same:
            a, b = _segments_key(full), _segments_key(res)
            k = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
            sys.stderr.write(f"[bench] rank {rank}: sharded result differs from the single-GPU result at segment {k} of {len(a)}/{len(b)}\n")
        # weak scaling beside it: every rank its own hour
        own = torch.from_numpy(make_audio(args.hours, 100 + rank)).to(device)
        transcribe(own, **kw)
        w_dev_s, _, _ = run(own, max(1, min(args.steps, 3)))
        extras["weak"] = {"value": args.hours * 3600.0 * world / (w_dev_s / max(1, min(args.steps, 3))), "unit": UNIT,
                          "ms_per_step": w_dev_s / max(1, min(args.steps, 3)) * 1e3, "scaling": "weak",
                          "note": "every rank transcribes its own hour (r01's line)"}
        del own

    line = None
    per_gpu = n_windows if (exact or not shard) else (n_windows + world - 1) // world
    if rank == 0:
        sm_mhz = clocks.get("sm_mhz") or 1900.0
        roof = kernel_rooflines(model, peaks, min(per_gpu, args.window_batch), with_logmel=(world == 1 and not args.no_extras))
        step = decode_step_probe(model, 1 if exact else min(per_gpu, args.window_batch), peaks)
        # the dominant kernel of the step is picked from the measured per-kernel shares of one eager step
        share = step["eager_kernel_share"]
        dom_name = next(iter(share))
        dom = roof.get(dom_name)  # profile labels of the library == roofline keys for the kernels that have one
        if dom is None:
            # a latency-bound phase dominates (small batches): report the whole step against its HBM floor instead
            dom = {"bound": "hbm", "achieved": step["algorithmic_bytes_per_step"] / (step["step_ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                   "unit": "GB/s", "frac": step["frac_of_hbm_floor"], "traffic": None}
            dom_label = f"whole decode step (dominant launch: {dom_name}, latency-bound)"
        else:
            dom_label = f"{dom_name} ({share[dom_name]:.0%} of the decode step)"
        cfg = workload_config(args.model, args.hours, args.window_batch)
        cfg.update({"mode": "exact sequential seek loop, batch 1" if exact else f"transcribe(window_batch={args.window_batch})",
                    "windows_per_gpu": per_gpu, "window_batch": 0 if exact else args.window_batch, "encoder_batch": args.encoder_batch,
                    "tokens_decoded_per_step": n_tokens, "l2": "working set (weights 3.1 GB + cross K/V 246 MB per window) >> 126 MB L2",
                    "parallelism": (f"{world} replica(s), one file each" if exact else
                                    (f"one hour's windows sharded over {world} GPU(s), host gather, no collective" if scaling == "strong"
                                     else f"{world} replica(s), one hour each, no collective"))})
        if exact:
            cfg["workload"] = (f"whisper-{args.model} (random-init) greedy transcription of {args.exact_seconds:g} s synthetic audio in the "
                               f"reference's exact sequential mode (batch 1), sample_len {SAMPLE_LEN}")
        line = {
            "metric": METRIC, "value": total_audio_s * args.steps / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": total_audio_s * args.steps / e2e_dev_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(audio_host.numel() * 4) * world,
                    "d2h_bytes_per_step": int(n_windows * (456 * 4 + 12)), "wall_value": total_audio_s * args.steps / e2e_wall_s},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {k: dom[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")},
            "roofline_kernel": dom_label,
            "roofline_peak_source": peaks["source"],
            "decode_step": step,
            "rooflines": roof,
            "wall_value": total_audio_s * args.steps / wall_s,
            **extras,
        }
        if world == 1 and not exact and not args.no_extras:
            line["exact_mode"] = exact_probe(model, args, transcribe)
            line["configs"] = other_configs(peaks, device, sm_mhz)
        # the CPU baseline is timed on rank 0 at N = 1 only (torchrun pins OMP_NUM_THREADS=1 for N > 1)
        model.release_sessions()
        line["cpu_baseline"] = cpu_baseline(args, weights) if (not args.no_cpu_baseline and world == 1) else None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def exact_probe(model, args, transcribe):
    """The default of transcribe() / the CLI (what `./run` hits): sequential seek loop, batch 1."""
    import torch

    seconds = args.exact_seconds
    audio = torch.from_numpy(make_audio(seconds / 3600.0, 7)).to(model.device)
    kw = dict(model=model, temperature=0.0, condition_on_previous_text=False, language="en", window_batch=0)
    transcribe(audio[: 16000 * 31], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = transcribe(audio, **kw)
    e1.record()
    torch.cuda.synchronize()
    s = e0.elapsed_time(e1) * 1e-3
    n_tok = sum(len(x["tokens"]) for x in r["segments"])
    out = {"value": seconds / s, "unit": UNIT, "audio_s": seconds, "device_s": s, "tokens": n_tok,
           "ms_per_token": s / max(n_tok, 1) * 1e3, "mode": "exact sequential seek loop, batch 1, temperature 0"}
    # the same exact loop over three files at once (transcribe_many: what the CLI does with several `audio` arguments)
    from whisper_mlx_b200 import transcribe_many as tm

    third = seconds / 3.0
    files = [torch.from_numpy(make_audio(third / 3600.0, 8 + i)).to(model.device) for i in range(3)]
    kw.pop("model")
    tm([x[: 16000 * 31] for x in files], model=model, **kw)
    torch.cuda.synchronize()
    e0.record()
    res = tm(files, model=model, **kw)
    e1.record()
    torch.cuda.synchronize()
    if all(isinstance(r, dict) for r in res):
        s3 = e0.elapsed_time(e1) * 1e-3
        out["three_files_lockstep"] = {"value": 3 * third / s3, "unit": UNIT, "audio_s": 3 * third, "device_s": s3,
                                       "mode": "transcribe_many: the exact loop of three files, decoder steps in shared batches"}
    return out


def other_configs(peaks, device, sm_mhz):
    """BASELINE configs[2] and [4]: decode steps of whisper-small at batch 64 and large-v3-turbo at batch 256."""
    import torch

    out = {}
    for key, name, batch in (("3", "small", 64), ("5", "large-v3-turbo", 256)):
        m, w = build_model(name, 0, device)
        r = decode_step_probe(m, batch, peaks)
        r["model"] = name
        r.pop("eager_kernel_share", None)
        out[key] = r
        m.release_sessions()
        del m, w
        torch.cuda.empty_cache()
    return out


def run_config(args):
    """`--config 2|3|5`: one line for another BASELINE configuration (N = 1)."""
    import torch

    peaks = load_peaks()
    torch.cuda.set_device(0)
    sampler = ClockSampler(0)
    sampler.start()
    if args.config == 2:
        roof = logmel_rooflines(peaks, "cuda:0")
        clocks = sampler.stop()
        r = roof["logmel_128"]
        line = {"metric": "log-mel frames/s vs HBM roofline", "value": r["frames_per_s"], "unit": "frames/s", "n_gpus": 1,
                "steps": 5, "warmup": 1, "ms_per_step": r["launch_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "BASELINE configs[1]: log-mel front-end, batch 1024 x 30 s windows, 128 mel bins (80 under rooflines), "
                                       "f32 in / f32 out, per-window clamp", "l2": "inputs 1.97 GB >> 126 MB L2"},
                "roofline": {k: r[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}, "rooflines": roof, "clocks": clocks}
    else:
        name, batch = ("small", 64) if args.config == 3 else ("large-v3-turbo", 256)
        m, _ = build_model(name, 0, "cuda:0")
        r = decode_step_probe(m, batch, peaks, steps=args.steps * 10)
        clocks = sampler.stop()
        line = {"metric": f"whisper-{name} greedy decode tokens/s", "value": r["tokens_per_s"], "unit": "tokens/s", "n_gpus": 1,
                "steps": args.steps * 10, "warmup": args.warmup, "ms_per_step": r["step_ms"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[{args.config - 1}]: whisper-{name} greedy decode, batch {batch} x 30 s windows, one "
                                       "single-token step (CUDA-graph replay) per bench step", "l2": "cross K/V >> 126 MB L2"},
                "roofline": {"bound": "hbm", "achieved": r["algorithmic_bytes_per_step"] / (r["step_ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": r["frac_of_hbm_floor"], "traffic": None},
                "decode_step": r, "clocks": clocks}
    print(json.dumps(line), flush=True)


# ======================================================================================== CPU arm
class CpuWindow:
    """One 30 s window through the oracle (PyTorch-CPU fp32 restatement of the reference algorithm), steppable:
    `start()` = log-mel + encoder + the prompt step, `advance(n)` = n greedy single-token steps."""

    def __init__(self, weights_f32, dims_dict, threads: int, seed: int = 7):
        import torch
        from oracle import model as OM
        from oracle.tokens import TokenIds
        from tools import synth

        torch.set_num_threads(threads)
        self.w = weights_f32
        self.dims = OM.ModelDimensions(**dims_dict)
        self.ids = TokenIds(self.dims.n_vocab)
        self.x = synth.white_noise(480000, seed)
        self.initial = list(self.ids.sot_sequence("en"))
        self.tokens = np.array([self.initial], dtype=np.int64)
        self.sum_lp = np.zeros(1, dtype=np.float32)
        self.cache = None
        self.xa = None
        self.steps_done = 0

    def start(self):
        import torch
        from oracle import audio as OA, model as OM

        t0 = time.perf_counter()
        mel = torch.from_numpy(OA.log_mel_spectrogram(self.x, self.dims.n_mels))[None]
        t1 = time.perf_counter()
        self.xa = OM.encoder_forward(self.w, self.dims, mel)
        t2 = time.perf_counter()
        return {"logmel_s": t1 - t0, "encoder_s": t2 - t1}

    def advance(self, n: int) -> float:
        import torch
        from oracle import decoding as OD, model as OM

        t0 = time.perf_counter()
        with torch.no_grad():
            for _ in range(n):
                inp = self.tokens if self.cache is None else self.tokens[:, -1:]
                logits, self.cache = OM.decoder_forward(self.w, self.dims, torch.from_numpy(inp), self.xa, self.cache)
                step = logits[:, -1].float().numpy().copy()
                OD.filter_logits(step, self.tokens, len(self.initial), self.ids, self.ids.suppress_set())
                self.tokens, _ = OD.greedy_update(self.tokens, step, self.sum_lp, self.ids.eot)
                self.steps_done += 1
        return time.perf_counter() - t0


def _cpu_weights(args, weights=None):
    import torch
    from tools import synth

    dims = synth.DIMS[args.model]
    if weights is None:
        weights = dict(synth.random_weights(dims, 0, device="cpu"))
    return dims, {k: v.detach().to("cpu", torch.float32) for k, v in weights.items()}


def cpu_baseline(args, weights=None):
    """The CPU port timed on the host cores beside the GPU run: one whole 30 s window, all 224 decode steps really run
    (with --cpu-sample-len n < 224 only n steps run and `value` is null: no extrapolated number is reported)."""
    threads = os.cpu_count() or 1
    dims, w32 = _cpu_weights(args, weights)
    win = CpuWindow(w32, dims, threads)
    r = win.start()
    n = min(args.cpu_sample_len, SAMPLE_LEN)
    r["decode_s"] = win.advance(n)
    r["decode_steps"] = n
    total = r["logmel_s"] + r["encoder_s"] + r["decode_s"]
    full = n == SAMPLE_LEN
    return {"value": 30.0 / total if full else None, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one 30 s window of the same workload through oracle/ (PyTorch-CPU fp32 restatement, not MLX; batch 1): log-mel "
                      f"{r['logmel_s']:.2f} s + encoder {r['encoder_s']:.2f} s + {n} greedy decode steps {r['decode_s']:.2f} s"
                      + ("" if full else f" (only {n} of {SAMPLE_LEN} steps: no RTFx reported)"),
            "detail": r}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.perf_counter()
    threads = os.cpu_count() or 1
    dims, w32 = _cpu_weights(args)
    # warm-up: the first steps of a scratch window (thread pool, page-in of the 6 GB of fp32 weights), untimed
    if args.warmup > 0:
        scratch = CpuWindow(w32, dims, threads, seed=8)
        import torch

        scratch.xa = torch.zeros(1, dims["n_audio_ctx"], dims["n_audio_state"])
        for _ in range(args.warmup):
            scratch.advance(2)
        del scratch
    # the K timed steps together are ONE full window: step 0 also carries the log-mel and the encoder
    K = max(1, args.steps)
    win = CpuWindow(w32, dims, threads)
    per_step = []
    detail = {}
    for i in range(K):
        t0 = time.perf_counter()
        if i == 0:
            detail = win.start()
        n = SAMPLE_LEN * (i + 1) // K - SAMPLE_LEN * i // K
        detail[f"decode_s_step{i}"] = win.advance(n)
        per_step.append(time.perf_counter() - t0)
    assert win.steps_done == SAMPLE_LEN
    total = float(sum(per_step))
    value = 30.0 / total
    cfg = workload_config(args.model, args.hours, args.window_batch)
    cfg.update({"mode": "oracle/ CPU restatement (mlx-whisper is not installable here), batch 1", "sample": "see cpu_baseline.sample"})
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": K, "warmup": args.warmup, "ms_per_step": total / K * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"the {K} timed steps together are ONE full 30 s window of the workload (a bounded sample of its "
                                   f"{cfg['windows']} windows): log-mel + encoder in step 0, all {SAMPLE_LEN} greedy decode steps split evenly "
                                   f"over the steps; nothing extrapolated; torch threads = {threads}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "detail": {k: round(v, 4) for k, v in detail.items()},
        "total_wall_s": time.perf_counter() - t_start,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="large-v3", choices=["tiny", "small", "large-v3", "large-v3-turbo", "micro"])
    ap.add_argument("--hours", type=float, default=1.0)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--mode", default="batched", choices=["batched", "exact"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5])
    ap.add_argument("--exact-seconds", type=float, default=300.0)
    ap.add_argument("--window-batch", type=int, default=120)
    ap.add_argument("--encoder-batch", type=int, default=40)
    ap.add_argument("--cpu-sample-len", type=int, default=SAMPLE_LEN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the exact-mode / configs 3, 5 / log-mel side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config != 4:
        run_config(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
