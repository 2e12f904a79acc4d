"""Command line: host mirror of `mlx_whisper/cli.py` (UPSTREAM; the console script invoked by
/root/reference/run:3-6 with `-f txt --output-name NAME --model ... --condition-on-previous-text False
--hallucination-silence-threshold 1`; restated in SURVEY.md A.7).  Same hyphenated flags and defaults.

    python -m whisper_mlx_b200 input.wav -f txt --output-name out --model /path/to/whisper-large-v3-mlx \
        --condition-on-previous-text False --hallucination-silence-threshold 1
"""
from __future__ import annotations

import argparse
import os
import traceback
import warnings

import numpy as np

from .tokenizer import LANGUAGES, TO_LANGUAGE_CODE
from .transcribe import transcribe
from .writers import get_writer


def optional_int(string):
    return None if string == "None" else int(string)


def optional_float(string):
    return None if string == "None" else float(string)


def str2bool(string):
    str2val = {"True": True, "False": False}
    if string in str2val:
        return str2val[string]
    raise ValueError(f"Expected one of {set(str2val.keys())}, got {string}")


def build_parser():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("audio", nargs="+", type=str, help="Audio file(s) to transcribe")
    parser.add_argument("--model", default="mlx-community/whisper-tiny", type=str, help="The model directory or hugging face repo")
    parser.add_argument("--output-name", type=str, default=None, help="The name of transcription/translation output files before --output-format extensions")
    parser.add_argument("--output-dir", "-o", type=str, default=".", help="Directory to save the outputs")
    parser.add_argument("--output-format", "-f", type=str, default="txt", choices=["txt", "vtt", "srt", "tsv", "json", "all"], help="Format of the output file")
    parser.add_argument("--verbose", type=str2bool, default=True, help="Whether to print out progress and debug messages")
    parser.add_argument("--task", type=str, default="transcribe", choices=["transcribe", "translate"], help="Perform speech recognition ('transcribe') or speech translation ('translate')")
    parser.add_argument("--language", type=str, default=None, choices=sorted(LANGUAGES.keys()) + sorted([k.title() for k in TO_LANGUAGE_CODE.keys()]), help="Language spoken in the audio, specify None to auto-detect")
    parser.add_argument("--temperature", type=float, default=0, help="Temperature for sampling")
    parser.add_argument("--best-of", type=optional_int, default=5, help="Number of candidates when sampling with non-zero temperature")
    parser.add_argument("--patience", type=float, default=None, help="Optional patience value to use in beam decoding")
    parser.add_argument("--length-penalty", type=float, default=None, help="Optional token length penalty coefficient (alpha)")
    parser.add_argument("--suppress-tokens", type=str, default="-1", help="Comma-separated list of token ids to suppress during sampling; '-1' will suppress most special characters except common punctuations")
    parser.add_argument("--initial-prompt", type=str, default=None, help="Optional text to provide as a prompt for the first window")
    parser.add_argument("--condition-on-previous-text", type=str2bool, default=True, help="If True, provide the previous output of the model as a prompt for the next window")
    parser.add_argument("--fp16", type=str2bool, default=True, help="Whether to perform inference in 16-bit storage")
    parser.add_argument("--compression-ratio-threshold", type=optional_float, default=2.4, help="If the gzip compression ratio is higher than this value, treat the decoding as failed")
    parser.add_argument("--logprob-threshold", type=optional_float, default=-1.0, help="If the average log probability is lower than this value, treat the decoding as failed")
    parser.add_argument("--no-speech-threshold", type=optional_float, default=0.6, help="If the probability of the token is higher than this value the segment is considered silence")
    parser.add_argument("--word-timestamps", type=str2bool, default=False, help="Extract word-level timestamps using the cross-attention pattern")
    parser.add_argument("--prepend-punctuations", type=str, default="\"'“¿([{-", help="If word-timestamps is True, merge these punctuation symbols with the next word")
    parser.add_argument("--append-punctuations", type=str, default="\"'.。,，!！?？:：”)]}、", help="If word_timestamps is True, merge these punctuation symbols with the previous word")
    parser.add_argument("--highlight-words", type=str2bool, default=False, help="(requires --word_timestamps True) underline each word as it is spoken in srt and vtt")
    parser.add_argument("--max-line-width", type=int, default=None, help="(requires --word_timestamps True) the maximum number of characters in a line before breaking the line")
    parser.add_argument("--max-line-count", type=int, default=None, help="(requires --word_timestamps True) the maximum number of lines in a segment")
    parser.add_argument("--max-words-per-line", type=int, default=None, help="(requires --word_timestamps True, no effect with --max_line_width) the maximum number of words in a segment")
    parser.add_argument("--hallucination-silence-threshold", type=optional_float, help="(requires --word_timestamps True) skip silent periods longer than this threshold (in seconds) when a possible hallucination is detected")
    parser.add_argument("--clip-timestamps", type=str, default="0", help="Comma-separated list start,end,start,end,... timestamps (in seconds) of clips to process")
    parser.add_argument("--temperature-increment-on-fallback", type=optional_float, default=0.2, help="Temperature to increase when falling back when the decoding fails to meet either of the thresholds below")
    parser.add_argument("--window-batch", type=int, default=None, help="(B200 extension) decode this many fixed 30 s windows together; needs --condition-on-previous-text False")
    return parser


def main(argv=None):
    parser = build_parser()
    args = vars(parser.parse_args(argv))
    if args["verbose"] is True:
        print(f"Args: {args}")

    path_or_hf_repo: str = args.pop("model")
    output_dir: str = args.pop("output_dir")
    output_format: str = args.pop("output_format")
    output_name: str = args.pop("output_name")
    os.makedirs(output_dir, exist_ok=True)

    writer = get_writer(output_format, output_dir)
    word_options = ["highlight_words", "max_line_count", "max_line_width", "max_words_per_line"]
    writer_args = {arg: args.pop(arg) for arg in word_options}
    if not args["word_timestamps"]:
        for k, v in writer_args.items():
            if v:
                parser.error(f"--{k.replace('_', '-')} requires --word-timestamps True")
    if writer_args["max_line_count"] and not writer_args["max_line_width"]:
        warnings.warn("--max-line-count has no effect without --max-line-width")
    if writer_args["max_words_per_line"] and writer_args["max_line_width"]:
        warnings.warn("--max-words-per-line has no effect with --max-line-width")

    temperature = args.pop("temperature")
    if (increment := args.pop("temperature_increment_on_fallback")) is not None:
        temperature = tuple(np.arange(temperature, 1.0 + 1e-6, increment))
    else:
        temperature = [temperature]

    if args["hallucination_silence_threshold"] is not None and not args["word_timestamps"]:
        warnings.warn("--hallucination-silence-threshold requires --word-timestamps True; it has no effect")

    audio_paths = args.pop("audio")
    lockstep = len(audio_paths) > 1 and not (args.get("window_batch") or int(os.environ.get("B200W_WINDOW_BATCH", "0")))
    if lockstep:
        # several files in the exact (sequential) mode: their decoder steps share batches (transcribe_many)
        from .transcribe import transcribe_many

        # (their segments would interleave on the terminal: each file's transcript is printed when all are done)
        verbose = args.get("verbose")
        outcomes = transcribe_many(audio_paths, path_or_hf_repo=path_or_hf_repo, temperature=temperature,
                                   **{**args, "verbose": None if verbose else verbose})
        if verbose:
            for path, r in zip(audio_paths, outcomes):
                if isinstance(r, dict):
                    print(f"== {path}")
                    for seg in r["segments"]:
                        print(f"[{seg['start']:9.3f} --> {seg['end']:9.3f}] {seg['text']}")
    for k, audio_path in enumerate(audio_paths):
        try:
            if lockstep:
                result = outcomes[k]
                if isinstance(result, BaseException):
                    raise result
            else:
                result = transcribe(audio_path, path_or_hf_repo=path_or_hf_repo, temperature=temperature, **args)
            name = output_name or os.path.splitext(os.path.basename(audio_path))[0]
            writer(result, name, **writer_args)
        except Exception as e:  # noqa: BLE001 - the reference reports per file and carries on
            traceback.print_exc()
            print(f"Skipping {audio_path} due to {type(e).__name__}: {str(e)}")


if __name__ == "__main__":
    main()
