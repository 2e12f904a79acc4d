"""Synthetic BPE vocabulary in the `*.tiktoken` format of the reference's assets (test fixture generator).

`mlx_whisper/assets/multilingual.tiktoken` (50257 ranks) and `gpt2.tiktoken` (50256 ranks) are not in this image and
cannot be fetched.  The file format is public: one `base64(token bytes) rank` pair per line.  This module writes a
vocabulary of the same size and format -- the 256 single bytes in GPT-2's byte order (so that " " is id 220, the id
SuppressBlank relies on) followed by deterministic merges of existing tokens -- which exercises the tiktoken branch of
whisper-mlx_b200/tokenizer.py (file lookup, parsing, special-token table after the ranks, encode / decode,
non_speech_tokens, word splitting) offline.  Token ids differ from the real vocabulary; nothing here claims otherwise.
"""
from __future__ import annotations

import base64
import os
import random


def gpt2_byte_order():
    """The byte <-> rank order of the GPT-2 / Whisper vocabularies: printable bytes first, then the rest."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    rest = [b for b in range(256) if b not in bs]
    return bs + rest


def build_ranks(n_ranks: int, seed: int = 0):
    order = gpt2_byte_order()
    tokens = [bytes([b]) for b in order]
    seen = set(tokens)
    rng = random.Random(seed)
    # a few hundred hand-picked merges so that common English pieces exist as multi-byte tokens ...
    # (every merge is of two EXISTING tokens, so BPE can reach it)
    seeds = [b" t", b"he", b" a", b"in", b" th", b"er", b" the", b"re", b"on", b" s", b"at", b" w", b"en", b" o", b"it", b"is",
             b"an", b"or", b"es", b" b", b"ed", b" f", b"ing", b" p", b"ou", b" an", b"al", b"ar", b" to", b" m", b" of",
             b" in", b" d", b" h", b" and", b"ic", b"as", b"le", b" th", b"ion", b"om", b"ll", b"ent", b" n", b" l", b"st",
             b" re", b"ve", b" e", b"ro", b"ly", b" be", b" g", b" T", b"ct", b" S", b"id", b"ot", b" I", b"ut", b"et", b" A",
             b" is", b" on", b"im", b"am", b"ow", b"ay", b"ad", b"se", b" that", b" C", b"ig", b" for", b"ac", b" y", b"ver",
             b"ur", b" u", b"ld", b" st", b" M", b"'s", b" he", b" it", b"ation", b"ith", b"ir", b"ce", b" you", b"il", b" B",
             b" wh", b"ol", b" P", b" with", b" 1", b"ter", b"ch", b" as", b" we", b" (", b"nd", b"ill", b" D", b"if", b" 2",
             b"ag", b"ers", b"ke", b' "', b" -", b" '", b"--", b"((", b"))", b" [", b" ]", b"[[", b"]]", b"\xe2\x99", b"\xe2\x99\xaa",
             b"\xe2\x99\xaa\xe2\x99\xaa", b" \xe2\x99", b" \xe2\x99\xaa", b"\xe3\x80", b"\xe3\x80\x8c", b"\xe3\x80\x8d"]

    def add(tok: bytes) -> bool:
        if tok in seen or len(tokens) >= n_ranks:
            return False
        seen.add(tok)
        tokens.append(tok)
        return True

    def reachable(tok: bytes) -> bool:  # some split into two existing tokens
        return any(tok[:i] in seen and tok[i:] in seen for i in range(1, len(tok)))

    for s in seeds:
        if reachable(s):
            add(s)
    while len(tokens) < n_ranks:
        pool = min(len(tokens), 4096 + len(tokens) // 8)  # merges of (mostly) short, early tokens, like a trained BPE
        a, b = tokens[rng.randrange(pool)], tokens[rng.randrange(pool)]
        if len(a) + len(b) <= 12:
            add(a + b)
    return {tok: i for i, tok in enumerate(tokens)}


def write_vocab(directory: str, name: str = "multilingual", seed: int = 0) -> str:
    """Write `<directory>/<name>.tiktoken` (50257 ranks for "multilingual", 50256 for "gpt2"); returns the path."""
    n = 50257 if name == "multilingual" else 50256
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, f"{name}.tiktoken")
    if not os.path.exists(path):
        ranks = build_ranks(n, seed)
        with open(path + ".tmp", "w") as f:
            for tok, rank in sorted(ranks.items(), key=lambda kv: kv[1]):
                f.write(f"{base64.b64encode(tok).decode()} {rank}\n")
        os.replace(path + ".tmp", path)
    return path


if __name__ == "__main__":
    import sys

    print(write_vocab(sys.argv[1] if len(sys.argv) > 1 else ".", sys.argv[2] if len(sys.argv) > 2 else "multilingual"))
