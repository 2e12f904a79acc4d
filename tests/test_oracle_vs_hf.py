"""Pin the oracle against the independent `transformers` Whisper implementation installed in this image
(the reference holds no golden vectors, so a second executable source pins the
restatement).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import audio as OA
from oracle import decoding as OD
from oracle import model as OM
from oracle.tokens import LANGUAGE_CODES, NON_SPEECH_MULTILINGUAL, TokenIds
from tools import synth


@pytest.mark.parametrize("n_mels", [80, 128])
def test_mel_filters_equal_hf(n_mels):
    from transformers import WhisperFeatureExtractor

    fe = WhisperFeatureExtractor(feature_size=n_mels)
    ours = OA.mel_filters(n_mels)
    assert np.abs(ours - fe.mel_filters.T.astype(np.float32)).max() <= 1e-7
    # structure the CUDA kernel relies on: bins 0 and 200 carry no weight, <= 2 non-zeros per frequency row
    assert not ours[:, 0].any() and not ours[:, 200].any()
    assert (ours != 0).sum(0).max() <= 2
    assert (ours != 0).sum() == {80: 391, 128: 394}[n_mels]


@pytest.mark.parametrize("kind,n_mels", [("noise", 80), ("noise", 128), ("tones", 80), ("speech", 128), ("clip", 80)])
def test_log_mel_equals_hf(kind, n_mels):
    from transformers import WhisperFeatureExtractor

    fe = WhisperFeatureExtractor(feature_size=n_mels)
    x = synth.make_audio(kind, 480000, 1)
    hf = fe._np_extract_fbank_features(x[None], "cpu")[0].T  # HF is (n_mels, frames)
    ours = OA.log_mel_spectrogram(x, n_mels)
    assert ours.shape == (3000, n_mels)
    assert np.abs(ours - hf).max() <= 2e-5


def test_log_mel_padding_and_ragged():
    x = synth.white_noise(16000 * 3 + 77, 2)
    a = OA.log_mel_spectrogram(x, 80, padding=480000)
    assert a.shape == ((len(x) + 480000) // 160, 80)
    b = OA.log_mel_spectrogram(np.concatenate([x, np.zeros(480000, np.float32)]), 80)
    assert np.array_equal(a, b)


def test_model_forward_equals_hf():
    from tests.hf_bridge import build_hf

    dims_d = synth.DIMS["micro"]
    w = {k: v.float() for k, v in synth.random_weights(dims_d, 3)}
    dims = OM.ModelDimensions(**dims_d)
    hf = build_hf(dims_d, w)
    mel = torch.from_numpy(OA.log_mel_spectrogram(synth.white_noise(480000, 0), 80))[None]
    xa = OM.encoder_forward(w, dims, mel)
    toks = torch.tensor([[50258, 50259, 50359, 50364, 400, 500, 50400, 50400, 7]])
    logits, cache = OM.decoder_forward(w, dims, toks, xa)
    with torch.no_grad():
        enc = hf.model.encoder(mel.transpose(1, 2)).last_hidden_state
        out = hf(input_features=mel.transpose(1, 2), decoder_input_ids=toks).logits
    assert (enc - xa).abs().max().item() <= 2e-4
    assert (out - logits).abs().max().item() <= 2e-4
    # incremental decoding with the KV cache reproduces the full pass
    l0, c = OM.decoder_forward(w, dims, toks[:, :3], xa)
    for i in range(3, toks.shape[1]):
        li, c = OM.decoder_forward(w, dims, toks[:, i: i + 1], xa, c)
        assert (li[:, 0] - logits[:, i]).abs().max().item() <= 1e-4


def test_special_token_table():
    for n_vocab, tb in ((51865, 50364), (51866, 50365)):
        ids = TokenIds(n_vocab)
        assert (ids.eot, ids.sot, ids.timestamp_begin) == (50257, 50258, tb)
        assert ids.no_timestamps == tb - 1 and ids.no_speech == tb - 2
        assert ids.timestamp_begin + 1501 == n_vocab
    from transformers.models.whisper.configuration_whisper import NON_SPEECH_TOKENS_MULTI
    from transformers.models.whisper.tokenization_whisper import LANGUAGES

    assert list(NON_SPEECH_MULTILINGUAL) == [t for t in NON_SPEECH_TOKENS_MULTI if t < 50257]
    assert list(LANGUAGES.keys()) == LANGUAGE_CODES


def test_timestamp_rules_equal_hf():
    """oracle.apply_timestamp_rules == transformers' WhisperTimeStampLogitsProcessor on random histories."""
    from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor

    class Cfg:
        no_timestamps_token_id = 50363
        eos_token_id = 50257
        max_initial_timestamp_index = 50

    ids = TokenIds(51865)
    tb = ids.timestamp_begin
    proc = WhisperTimeStampLogitsProcessor(Cfg(), begin_index=3)
    rng = np.random.default_rng(0)
    histories = [[], [tb + 3], [tb + 3, 100], [tb + 3, 100, tb + 9], [tb + 3, 100, tb + 9, tb + 9], [tb + 3, 100, tb + 9, tb + 9, 5, 6],
                 [tb, tb], [tb + 1500], [tb + 7, 11, 12, 13]]
    for h in histories:
        for shift in (0.0, 8.0, -8.0):
            toks = np.array([[50258, 50259, 50359] + h], dtype=np.int64)
            logits = rng.standard_normal((1, 51865)).astype(np.float32) * 2
            logits[:, tb:] += shift
            ours = logits.copy()
            OD.apply_timestamp_rules(ours, toks, 3, ids, 50)
            theirs = proc(torch.from_numpy(toks), torch.from_numpy(logits.copy())).numpy()
            assert np.array_equal(np.isneginf(ours), np.isneginf(theirs)), h
            assert np.array_equal(ours[~np.isneginf(ours)], theirs[~np.isneginf(theirs)])


def test_alignment_primitives_match_hf():
    """oracle.timing's median filter and DTW (restating mlx_whisper/timing.py) against the independent implementations
    in transformers (`generation_whisper._median_filter`, `_dynamic_time_warping`): identical filter output and
    identical alignment paths on random cost matrices."""
    import numpy as np
    import torch
    from transformers.models.whisper import generation_whisper as G

    from oracle import timing as OT

    rng = np.random.default_rng(0)
    for shape in ((2, 5, 60), (1, 3, 9), (4, 2, 1500)):
        x = rng.standard_normal(shape).astype(np.float32)
        assert np.array_equal(OT.median_filter(x, 7), G._median_filter(torch.from_numpy(x), 7).numpy())
    for n, m in ((23, 180), (1, 40), (40, 41), (30, 500)):
        cost = rng.standard_normal((n, m)).astype(np.float32)
        ti, tj = OT.dtw(-cost)
        hi, hj = G._dynamic_time_warping(-cost.astype(np.float64))
        assert np.array_equal(ti, hi) and np.array_equal(tj, hj), (n, m)
        assert ti[0] == 0 and tj[0] == 0 and ti[-1] == n - 1 and tj[-1] == m - 1
        assert np.all(np.diff(ti) >= 0) and np.all(np.diff(tj) >= 0)
