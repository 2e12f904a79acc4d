"""Oracle: Whisper special-token table (test infrastructure only; see oracle/__init__.py).

Restates the id layout of `mlx_whisper/tokenizer.py` (UPSTREAM):
specials are appended after the 50257 BPE ranks in the order eot, sot, languages, translate,
transcribe, startoflm, startofprev, nospeech, notimestamps, 1501 timestamps.
Cross-pin: transformers/models/whisper/configuration_whisper.py:23-44 (non-speech ids).
"""
from __future__ import annotations

# first 82 entries of HF NON_SPEECH_TOKENS_MULTI (all < 50257); SURVEY.md Appendix B.2
NON_SPEECH_MULTILINGUAL = (
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542,
    873, 893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667,
    6585, 6647, 7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618,
    16553, 16604, 18362, 18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470,
    36865, 42863, 47425, 49870, 50254,
)

LANGUAGE_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi "
    "ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd "
    "gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue"
).split()


class TokenIds:
    def __init__(self, n_vocab: int):
        assert n_vocab >= 51865, "only multilingual vocabularies are on the reference path"
        self.n_vocab = n_vocab
        self.num_languages = n_vocab - 51765 - 1
        self.eot = 50257
        self.sot = 50258
        self.language_begin = 50259
        self.translate = self.language_begin + self.num_languages
        self.transcribe = self.translate + 1
        self.sot_lm = self.translate + 2
        self.sot_prev = self.translate + 3
        self.no_speech = self.translate + 4
        self.no_timestamps = self.translate + 5
        self.timestamp_begin = self.translate + 6
        self.blank = 220  # encode(" ")
        assert self.timestamp_begin + 1501 == n_vocab

    def language_token(self, code: str) -> int:
        return self.language_begin + LANGUAGE_CODES[: self.num_languages].index(code)

    def sot_sequence(self, language: str = "en", task: str = "transcribe"):
        return (self.sot, self.language_token(language), self.transcribe if task == "transcribe" else self.translate)

    def suppress_set(self):
        """`_get_suppress_tokens` with suppress_tokens="-1" (SURVEY.md A.4)."""
        ids = set(NON_SPEECH_MULTILINGUAL)
        ids |= {self.transcribe, self.translate, self.sot, self.sot_prev, self.sot_lm, self.no_speech}
        return tuple(sorted(ids))


def surrogate_piece(token_id: int) -> bytes:
    """Offline stand-in for a BPE piece when `multilingual.tiktoken` is not on disk.

    The real vocabulary file (upstream assets/multilingual.tiktoken) is absent from this image
    (SURVEY.md section 7.2-8); text is then rendered with this deterministic surrogate so that
    `compression_ratio` and the segment/`text` plumbing stay exercisable.  Token ids -- not text --
    are what parity is judged on.
    """
    h = (token_id * 2654435761) & 0xFFFFFFFF
    n = 1 + (h >> 28) % 5
    out = bytearray(b" " if (h >> 27) & 1 else b"")
    for _ in range(n):
        h = (h * 1103515245 + 12345) & 0x7FFFFFFF
        out.append(97 + (h >> 16) % 26)
    return bytes(out)


def special_name(token_id: int, timestamp_begin: int) -> str:
    """Name of a special token below the timestamps, as tiktoken renders it (specials follow the 50257
    ranks in the order eot, sot, languages, translate, transcribe, startoflm, startofprev, nospeech, notimestamps)."""
    n_lang = timestamp_begin - 50259 - 6
    names = ["endoftext", "startoftranscript", *LANGUAGE_CODES[:n_lang], "translate", "transcribe", "startoflm", "startofprev",
             "nospeech", "notimestamps"]
    return f"<|{names[token_id - 50257]}|>"


def decode_text(token_ids, timestamp_begin: int) -> str:
    """`Tokenizer.decode`: drops ids >= timestamp_begin, joins pieces (surrogate vocabulary for the BPE ranks; special
    tokens below the timestamps -- a model can sample e.g. a language token -- are rendered by name like tiktoken does)."""
    out = []
    for t in token_ids:
        t = int(t)
        if t < 50257:
            out.append(surrogate_piece(t))
        elif t < timestamp_begin:
            out.append(special_name(t, timestamp_begin).encode())
    return b"".join(out).decode("utf-8")
