"""Tokenizer: host mirror of `mlx_whisper/tokenizer.py` (UPSTREAM; reached from /root/reference/run:3-6;
restated in SURVEY.md A.6 and Appendix B.2).

The special-token table is computed exactly as the reference does (specials appended after the BPE
ranks in a fixed order).  Text <-> ids needs the BPE vocabulary file (`multilingual.tiktoken` /
`gpt2.tiktoken`), which the reference ships as a package asset; it is looked up in, in order,
$B200W_TIKTOKEN_DIR, <this package>/assets/ and the model directory.  When it is not on disk the tokenizer
REFUSES to be built (a transcript rendered from a made-up vocabulary must never reach a .txt file or a
compression-ratio decision).  Only with B200W_ALLOW_SURROGATE=1 -- set by the test-suite and the benchmark,
which run random-init weights and compare token ids -- are ids rendered with a deterministic surrogate
vocabulary so that the text / compression-ratio plumbing stays exercisable; token ids are unaffected.
"""
from __future__ import annotations

import base64
import os
import string
from dataclasses import dataclass, field
from functools import cached_property, lru_cache
from typing import Dict, List, Optional, Tuple

LANGUAGES = {
    "en": "english", "zh": "chinese", "de": "german", "es": "spanish", "ru": "russian", "ko": "korean",
    "fr": "french", "ja": "japanese", "pt": "portuguese", "tr": "turkish", "pl": "polish", "ca": "catalan",
    "nl": "dutch", "ar": "arabic", "sv": "swedish", "it": "italian", "id": "indonesian", "hi": "hindi",
    "fi": "finnish", "vi": "vietnamese", "he": "hebrew", "uk": "ukrainian", "el": "greek", "ms": "malay",
    "cs": "czech", "ro": "romanian", "da": "danish", "hu": "hungarian", "ta": "tamil", "no": "norwegian",
    "th": "thai", "ur": "urdu", "hr": "croatian", "bg": "bulgarian", "lt": "lithuanian", "la": "latin",
    "mi": "maori", "ml": "malayalam", "cy": "welsh", "sk": "slovak", "te": "telugu", "fa": "persian",
    "lv": "latvian", "bn": "bengali", "sr": "serbian", "az": "azerbaijani", "sl": "slovenian", "kn": "kannada",
    "et": "estonian", "mk": "macedonian", "br": "breton", "eu": "basque", "is": "icelandic", "hy": "armenian",
    "ne": "nepali", "mn": "mongolian", "bs": "bosnian", "kk": "kazakh", "sq": "albanian", "sw": "swahili",
    "gl": "galician", "mr": "marathi", "pa": "punjabi", "si": "sinhala", "km": "khmer", "sn": "shona",
    "yo": "yoruba", "so": "somali", "af": "afrikaans", "oc": "occitan", "ka": "georgian", "be": "belarusian",
    "tg": "tajik", "sd": "sindhi", "gu": "gujarati", "am": "amharic", "yi": "yiddish", "lo": "lao", "uz": "uzbek",
    "fo": "faroese", "ht": "haitian creole", "ps": "pashto", "tk": "turkmen", "nn": "nynorsk", "mt": "maltese",
    "sa": "sanskrit", "lb": "luxembourgish", "my": "myanmar", "bo": "tibetan", "tl": "tagalog", "mg": "malagasy",
    "as": "assamese", "tt": "tatar", "haw": "hawaiian", "ln": "lingala", "ha": "hausa", "ba": "bashkir",
    "jw": "javanese", "su": "sundanese", "yue": "cantonese",
}

# language code lookup by name, with a few language aliases
TO_LANGUAGE_CODE = {
    **{language: code for code, language in LANGUAGES.items()},
    "burmese": "my", "valencian": "ca", "flemish": "nl", "haitian": "ht", "letzeburgesch": "lb", "pushto": "ps",
    "panjabi": "pa", "moldavian": "ro", "moldovan": "ro", "sinhalese": "si", "castilian": "es", "mandarin": "zh",
}

# ids that `non_speech_tokens` evaluates to on the multilingual vocabulary
_NON_SPEECH_MULTILINGUAL = (
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542,
    873, 893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667,
    6585, 6647, 7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618,
    16553, 16604, 18362, 18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470,
    36865, 42863, 47425, 49870, 50254,
)


def _surrogate_piece(token_id: int) -> bytes:
    h = (token_id * 2654435761) & 0xFFFFFFFF
    n = 1 + (h >> 28) % 5
    out = bytearray(b" " if (h >> 27) & 1 else b"")
    for _ in range(n):
        h = (h * 1103515245 + 12345) & 0x7FFFFFFF
        out.append(97 + (h >> 16) % 26)
    return bytes(out)


def _find_vocab(name: str, extra_dirs=()) -> Optional[str]:
    dirs = [os.environ.get("B200W_TIKTOKEN_DIR"), os.path.join(os.path.dirname(__file__), "assets"), *extra_dirs]
    for d in dirs:
        if d and os.path.exists(os.path.join(d, f"{name}.tiktoken")):
            return os.path.join(d, f"{name}.tiktoken")
    return None


class _Encoding:
    """BPE core: tiktoken when the vocabulary file exists, else the surrogate."""

    def __init__(self, name: str, num_languages: int, vocab_path: Optional[str]):
        self.name = name
        n_base = 50257 if name == "multilingual" else 50256
        specials = ["<|endoftext|>", "<|startoftranscript|>", *[f"<|{lang}|>" for lang in list(LANGUAGES.keys())[:num_languages]],
                    "<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>",
                    "<|notimestamps|>", *[f"<|{i * 0.02:.2f}|>" for i in range(1501)]]
        self.special_tokens: Dict[str, int] = {tok: n_base + i for i, tok in enumerate(specials)}
        self.n_vocab = n_base + len(specials)
        self.eot_token = self.special_tokens["<|endoftext|>"]
        self._special_by_id = {v: k for k, v in self.special_tokens.items()}
        self._tk = None
        if vocab_path is not None:
            import tiktoken

            with open(vocab_path) as f:
                ranks = {base64.b64decode(tok): int(rank) for tok, rank in (line.split() for line in f if line)}
            self._tk = tiktoken.Encoding(
                name=os.path.basename(vocab_path), explicit_n_vocab=self.n_vocab,
                pat_str=r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""",
                mergeable_ranks=ranks, special_tokens=self.special_tokens)

    @property
    def has_vocab(self) -> bool:
        return self._tk is not None

    def encode(self, text: str) -> List[int]:
        if self._tk is not None:
            return self._tk.encode(text)
        if text == " ":
            return [220]  # the one encoding the decode loop itself needs (SuppressBlank)
        raise RuntimeError("encoding text needs the BPE vocabulary file (multilingual.tiktoken); "
                           "set B200W_TIKTOKEN_DIR or pass token ids")

    def decode(self, ids: List[int]) -> str:
        if self._tk is not None:
            return self._tk.decode(ids)
        table = self._piece_table()
        return b"".join(table[t] for t in ids).decode("utf-8", errors="replace")

    def _piece_table(self) -> List[bytes]:
        """id -> bytes for the surrogate vocabulary, built once (decode is on the per-window host path)."""
        if getattr(self, "_table", None) is None:
            n_base = min(self.special_tokens.values())
            table = [_surrogate_piece(t) for t in range(n_base)]
            table.extend(self._special_by_id[t].encode() for t in range(n_base, self.n_vocab))
            self._table = table
        return self._table


@dataclass
class Tokenizer:
    """A thin wrapper around the BPE encoding providing quick access to special tokens."""

    encoding: _Encoding
    num_languages: int
    language: Optional[str] = None
    task: Optional[str] = None
    sot_sequence: Tuple[int] = ()
    special_tokens: Dict[str, int] = field(default_factory=dict)

    def __post_init__(self):
        self.special_tokens = dict(self.encoding.special_tokens)
        sot = self.special_tokens["<|startoftranscript|>"]
        translate = self.special_tokens["<|translate|>"]
        transcribe = self.special_tokens["<|transcribe|>"]
        langs = tuple(list(LANGUAGES.keys())[: self.num_languages])
        sot_sequence = [sot]
        if self.language is not None:
            sot_sequence.append(sot + 1 + langs.index(self.language))
        if self.task is not None:
            sot_sequence.append(transcribe if self.task == "transcribe" else translate)
        self.sot_sequence = tuple(sot_sequence)

    def encode(self, text, **kwargs):
        return self.encoding.encode(text, **kwargs)

    def decode(self, token_ids: List[int], **kwargs) -> str:
        token_ids = [int(t) for t in token_ids if t < self.timestamp_begin]
        return self.encoding.decode(token_ids, **kwargs)

    def decode_with_timestamps(self, token_ids: List[int], **kwargs) -> str:
        """Timestamp tokens are above other special tokens' id range and are rendered as "<|1.08|>"."""
        return self.encoding.decode([int(t) for t in token_ids], **kwargs)

    @cached_property
    def eot(self) -> int:
        return self.encoding.eot_token

    @cached_property
    def transcribe(self) -> int:
        return self.special_tokens["<|transcribe|>"]

    @cached_property
    def translate(self) -> int:
        return self.special_tokens["<|translate|>"]

    @cached_property
    def sot(self) -> int:
        return self.special_tokens["<|startoftranscript|>"]

    @cached_property
    def sot_lm(self) -> int:
        return self.special_tokens["<|startoflm|>"]

    @cached_property
    def sot_prev(self) -> int:
        return self.special_tokens["<|startofprev|>"]

    @cached_property
    def no_speech(self) -> int:
        return self.special_tokens["<|nospeech|>"]

    @cached_property
    def no_timestamps(self) -> int:
        return self.special_tokens["<|notimestamps|>"]

    @cached_property
    def timestamp_begin(self) -> int:
        return self.special_tokens["<|0.00|>"]

    @cached_property
    def language_token(self) -> int:
        """Returns the token id corresponding to the value of the `language` field"""
        if self.language is None:
            raise ValueError("This tokenizer does not have language token configured")
        return self.to_language_token(self.language)

    def to_language_token(self, language):
        if token := self.special_tokens.get(f"<|{language}|>", None):
            return token
        raise KeyError(f"Language {language} not found in tokenizer.")

    def language_code(self, token: int) -> str:
        return self.all_language_codes[self.all_language_tokens.index(int(token))]

    @cached_property
    def all_language_tokens(self) -> Tuple[int]:
        result = []
        for token, token_id in self.special_tokens.items():
            if token.strip("<|>") in LANGUAGES:
                result.append(token_id)
        return tuple(result)[: self.num_languages]

    @cached_property
    def all_language_codes(self) -> Tuple[str]:
        inv = {v: k for k, v in self.special_tokens.items()}
        return tuple(inv[t].strip("<|>") for t in self.all_language_tokens)

    @cached_property
    def sot_sequence_including_notimestamps(self) -> Tuple[int]:
        return tuple(list(self.sot_sequence) + [self.no_timestamps])

    @cached_property
    def non_speech_tokens(self) -> Tuple[int]:
        """Tokens to suppress so that speaker tags / non-speech annotations ("[laughing]", "♪♪♪") are not sampled."""
        if not self.encoding.has_vocab:
            if self.encoding.name != "multilingual":
                raise RuntimeError("non_speech_tokens of the English-only vocabulary needs gpt2.tiktoken on disk")
            return _NON_SPEECH_MULTILINGUAL
        symbols = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』')
        symbols += "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        # symbols that may be a single token or multiple tokens depending on the tokenizer
        miscellaneous = set("♩♪♫♬♭♮♯")
        assert all(0x2640 <= ord(c) <= 0x267F for c in miscellaneous)
        # allow hyphens "-" and single quotes "'" between words, but not at the beginning of a word
        result = {self.encoding.encode(" -")[0], self.encoding.encode(" '")[0]}
        for symbol in symbols + list(miscellaneous):
            for tokens in [self.encoding.encode(symbol), self.encoding.encode(" " + symbol)]:
                if len(tokens) == 1 or symbol in miscellaneous:
                    result.add(tokens[0])
        return tuple(sorted(result))

    def split_to_word_tokens(self, tokens: List[int]):
        if self.language in {"zh", "ja", "th", "lo", "my", "yue"}:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)

    def split_tokens_on_unicode(self, tokens: List[int]):
        decoded_full = self.decode_with_timestamps(tokens)
        replacement_char = "�"
        words, word_tokens, current_tokens = [], [], []
        unicode_offset = 0
        for token in tokens:
            current_tokens.append(token)
            decoded = self.decode_with_timestamps(current_tokens)
            if replacement_char not in decoded or decoded_full[unicode_offset + decoded.index(replacement_char)] == replacement_char:
                words.append(decoded)
                word_tokens.append(current_tokens)
                current_tokens = []
                unicode_offset += len(decoded)
        return words, word_tokens

    def split_tokens_on_spaces(self, tokens: List[int]):
        subwords, subword_tokens_list = self.split_tokens_on_unicode(tokens)
        words, word_tokens = [], []
        for subword, subword_tokens in zip(subwords, subword_tokens_list):
            special = subword_tokens[0] >= self.eot
            with_space = subword.startswith(" ")
            punctuation = subword.strip() in string.punctuation
            if special or with_space or punctuation or len(words) == 0:
                words.append(subword)
                word_tokens.append(subword_tokens)
            else:
                words[-1] = words[-1] + subword
                word_tokens[-1].extend(subword_tokens)
        return words, word_tokens


def surrogate_allowed() -> bool:
    return os.environ.get("B200W_ALLOW_SURROGATE", "0") == "1"


@lru_cache(maxsize=None)
def _get_encoding(name: str, num_languages: int, vocab_path: Optional[str]) -> _Encoding:
    return _Encoding(name, num_languages, vocab_path)


def get_encoding(name: str = "gpt2", num_languages: int = 99, vocab_dir: Optional[str] = None) -> _Encoding:
    path = _find_vocab(name, (vocab_dir,) if vocab_dir else ())
    if path is None and not surrogate_allowed():
        where = [os.environ.get("B200W_TIKTOKEN_DIR") or "$B200W_TIKTOKEN_DIR (unset)",
                 os.path.join(os.path.dirname(__file__), "assets"), vocab_dir or "<model directory> (not given)"]
        raise FileNotFoundError(
            f"{name}.tiktoken (the BPE vocabulary mlx_whisper ships under mlx_whisper/assets/) was not found in "
            f"{', '.join(where)}: text cannot be produced without it.  Copy the file into one of these directories.  "
            "(B200W_ALLOW_SURROGATE=1 renders ids with a made-up vocabulary -- for tests and benchmarks only.)")
    return _get_encoding(name, num_languages, path)


def get_tokenizer(multilingual: bool, *, num_languages: int = 99, language: Optional[str] = None,
                  task: Optional[str] = None, vocab_dir: Optional[str] = None) -> Tokenizer:
    if language is not None:
        language = language.lower()
        if language not in LANGUAGES:
            if language in TO_LANGUAGE_CODE:
                language = TO_LANGUAGE_CODE[language]
            else:
                raise ValueError(f"Unsupported language: {language}")
    if multilingual:
        encoding_name = "multilingual"
        language = language or "en"
        task = task or "transcribe"
    else:
        encoding_name = "gpt2"
        language = None
        task = None
    encoding = get_encoding(name=encoding_name, num_languages=num_languages, vocab_dir=vocab_dir)
    key = (id(encoding), num_languages, language, task)
    if key not in _TOKENIZERS:
        _TOKENIZERS[key] = Tokenizer(encoding=encoding, num_languages=num_languages, language=language, task=task)
    return _TOKENIZERS[key]


_TOKENIZERS: Dict[tuple, Tokenizer] = {}
