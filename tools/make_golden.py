"""Generates tests/golden/*.npz|json from the CPU oracle (which tests/test_oracle_vs_hf.py pins against the
independent transformers implementation).  Run once in the build container; the fixtures travel with the
repo because neither /root/reference nor transformers' weights exist on the GPU box.

    python tools/make_golden.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import audio as OA, decoding as OD, model as OM, transcribe as OT  # noqa: E402
from oracle.tokens import TokenIds  # noqa: E402
from tools import synth  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def logmel():
    d = {}
    for kind in ("noise", "tones", "speech", "click", "clip"):
        x = synth.make_audio(kind, 16000 + 37, 11)
        for n_mels in (80, 128):
            d[f"{kind}_{n_mels}"] = OA.log_mel_spectrogram(x, n_mels)
    x = synth.white_noise(4000, 12)
    d["padded_80"] = OA.log_mel_spectrogram(x, 80, padding=1600)
    np.savez_compressed(os.path.join(OUT, "logmel.npz"), **d)


def model():
    dims_d = synth.DIMS["micro"]
    w = {k: v.float() for k, v in synth.random_weights(dims_d, 0)}
    dims = OM.ModelDimensions(**dims_d)
    ids = TokenIds(dims.n_vocab)
    tb = ids.timestamp_begin
    x = synth.make_audio("speech", 480000, 21)
    mel = torch.from_numpy(OA.log_mel_spectrogram(x, dims.n_mels))[None]
    rows = [0, 1, 700, 1499]
    seq = list(ids.sot_sequence("en")) + [tb + 5, 300, 4000, tb + 80, tb + 80, 900, tb + 200]
    toks = torch.tensor([seq])
    d = {"rows": np.array(rows), "tokens": np.array(seq)}
    for pol in ("fp32", "bf16"):
        xa = OM.encoder_forward(w, dims, mel, policy=pol)
        d[f"enc_{pol}"] = xa[0, rows].numpy()
        logits, _ = OM.decoder_forward(w, dims, toks, xa, policy=pol)
        top = logits[0].topk(8, dim=-1)
        d[f"top_ids_{pol}"] = top.indices.numpy()
        d[f"top_vals_{pol}"] = top.values.numpy()
        d[f"lse_{pol}"] = torch.logsumexp(logits[0], -1).numpy()
        res = OD.decode(w, dims, mel, language="en", sample_len=24, policy=pol, audio_features=xa)[0]
        d[f"greedy_{pol}"] = np.array(res.tokens)
        d[f"greedy_meta_{pol}"] = np.array([res.avg_logprob, res.no_speech_prob])
        d[f"greedy_margins_{pol}"] = np.array(res.margins, dtype=np.float32)  # top-1 / top-2 gap at every sampled token
    np.savez_compressed(os.path.join(OUT, "model_micro.npz"), **d)


def transcribe():
    dims_d = synth.DIMS["micro"]
    w = {k: v.float() for k, v in synth.random_weights(dims_d, 0)}
    dims = OM.ModelDimensions(**dims_d)
    audio = synth.long_audio(75.0, 3)
    out = {}
    for mode, fixed in (("exact", False), ("fixed", True)):
        windows = []

        def decode_fn(seek, size, segment, prompt, temperature):
            res = OD.decode(w, dims, segment, language="en", temperature=temperature, prompt=prompt, policy="bf16", sample_len=24)[0]
            windows.append({"seek": int(seek), "size": int(size), "tokens": res.tokens, "margins": [round(x, 5) for x in res.margins]})
            return res

        r = OT.transcribe(w, dims, audio, temperature=0.0, condition_on_previous_text=False, language="en", sample_len=24,
                          policy="bf16", fixed_windows=fixed, decode_fn=decode_fn)
        out[mode] = {"language": r["language"], "text": r["text"],
                     "segments": [{k: s[k] for k in ("id", "seek", "start", "end", "tokens", "text")} for s in r["segments"]],
                     "windows": windows}  # decoding order, with the top-1 / top-2 gap of every sampled token
    r = OT.transcribe(w, dims, audio[: 16000 * 31], temperature=0.0, condition_on_previous_text=False, sample_len=8)
    out["language_detect"] = r["language"]
    with open(os.path.join(OUT, "transcribe_micro.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    logmel()
    model()
    transcribe()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
