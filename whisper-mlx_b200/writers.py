"""Result writers: host mirror of `mlx_whisper/writers.py` (UPSTREAM; `-f txt --output-name` in
/root/reference/run:3; restated in SURVEY.md A.7)."""
from __future__ import annotations

import json
import os
import re
from typing import Callable, Optional, TextIO


def format_timestamp(seconds: float, always_include_hours: bool = False, decimal_marker: str = "."):
    assert seconds >= 0, "non-negative timestamp expected"
    milliseconds = round(seconds * 1000.0)
    hours = milliseconds // 3_600_000
    milliseconds -= hours * 3_600_000
    minutes = milliseconds // 60_000
    milliseconds -= minutes * 60_000
    seconds = milliseconds // 1_000
    milliseconds -= seconds * 1_000
    hours_marker = f"{hours:02d}:" if always_include_hours or hours > 0 else ""
    return f"{hours_marker}{minutes:02d}:{seconds:02d}{decimal_marker}{milliseconds:03d}"


class ResultWriter:
    extension: str

    def __init__(self, output_dir: str):
        self.output_dir = output_dir

    def __call__(self, result: dict, output_name: str, options: Optional[dict] = None, **kwargs):
        output_path = os.path.join(self.output_dir, output_name + "." + self.extension)
        with open(output_path, "w", encoding="utf-8") as f:
            self.write_result(result, file=f, options=options, **kwargs)

    def write_result(self, result: dict, file: TextIO, options: Optional[dict] = None, **kwargs):
        raise NotImplementedError


class WriteTXT(ResultWriter):
    extension: str = "txt"

    def write_result(self, result: dict, file: TextIO, options: Optional[dict] = None, **kwargs):
        for segment in result["segments"]:
            print(segment["text"].strip(), file=file, flush=True)


class SubtitlesWriter(ResultWriter):
    """vtt / srt cues.  Without word timings a cue is a segment.  With them (`--word-timestamps True`) the cue layout
    follows the reference's options: `max_line_width` wraps lines at that many characters, `max_line_count` closes a cue
    after that many lines (and at pauses longer than 3 s), `max_words_per_line` (only without `max_line_width`) cuts a
    segment into runs of that many words, `highlight_words` emits one cue per word with that word underlined."""

    always_include_hours: bool
    decimal_marker: str

    def iterate_result(self, result: dict, options: Optional[dict] = None, *, max_line_width: Optional[int] = None,
                       max_line_count: Optional[int] = None, highlight_words: bool = False,
                       max_words_per_line: Optional[int] = None):
        options = options or {}
        max_line_width = max_line_width or options.get("max_line_width")
        max_line_count = max_line_count or options.get("max_line_count")
        highlight_words = highlight_words or options.get("highlight_words", False)
        max_words_per_line = max_words_per_line or options.get("max_words_per_line")
        segments = result["segments"]
        if not (len(segments) > 0 and "words" in segments[0]):
            if any([max_line_width, max_line_count, highlight_words, max_words_per_line]):
                raise ValueError("highlight_words / max_line_width / max_line_count / max_words_per_line need word-level "
                                 "timestamps: transcribe with word_timestamps=True")
            for segment in segments:
                yield (self.format_timestamp(segment["start"]), self.format_timestamp(segment["end"]),
                       segment["text"].strip().replace("-->", "->"))
            return

        # cues keep to segment boundaries unless both a width and a line count are given
        keep_segments = max_line_count is None or max_line_width is None
        width = max_line_width or 1000
        per_line = max_words_per_line or 1000

        def cues():
            cue, line_len, n_lines = [], 0, 1
            last_start = segments[0]["start"]
            for segment in segments:
                words = segment["words"]
                for c0 in range(0, len(words), per_line):
                    for i, original in enumerate(words[c0: c0 + per_line]):
                        timing = dict(original)
                        long_pause = not keep_segments and timing["start"] - last_start > 3.0
                        fits = line_len + len(timing["word"]) <= width
                        new_run = i == 0 and len(cue) > 0 and keep_segments
                        if line_len > 0 and fits and not long_pause and not new_run:
                            line_len += len(timing["word"])  # stays on the current line
                        else:
                            timing["word"] = timing["word"].strip()
                            if (len(cue) > 0 and max_line_count is not None and (long_pause or n_lines >= max_line_count)) or new_run:
                                yield cue  # the cue is full (or the segment / word run ended)
                                cue, n_lines = [], 1
                            elif line_len > 0:
                                n_lines += 1  # wrap inside the cue
                                timing["word"] = "\n" + timing["word"]
                            line_len = len(timing["word"].strip())
                        cue.append(timing)
                        last_start = timing["start"]
            if cue:
                yield cue

        for cue in cues():
            cue_start = self.format_timestamp(cue[0]["start"])
            cue_end = self.format_timestamp(cue[-1]["end"])
            pieces = [w["word"] for w in cue]
            text = "".join(pieces).replace("-->", "->")
            if not highlight_words:
                yield cue_start, cue_end, text
                continue
            last = cue_start
            for i, w in enumerate(cue):
                start, end = self.format_timestamp(w["start"]), self.format_timestamp(w["end"])
                if last != start:
                    yield last, start, text  # the gap before this word: nothing underlined
                marked = [re.sub(r"^(\s*)(.*)$", r"\1<u>\2</u>", p) if j == i else p for j, p in enumerate(pieces)]
                yield start, end, "".join(marked).replace("-->", "->")
                last = end

    def format_timestamp(self, seconds: float):
        return format_timestamp(seconds=seconds, always_include_hours=self.always_include_hours,
                                decimal_marker=self.decimal_marker)


class WriteVTT(SubtitlesWriter):
    extension: str = "vtt"
    always_include_hours: bool = False
    decimal_marker: str = "."

    def write_result(self, result: dict, file: TextIO, options: Optional[dict] = None, **kwargs):
        print("WEBVTT\n", file=file)
        for start, end, text in self.iterate_result(result, options, **kwargs):
            print(f"{start} --> {end}\n{text}\n", file=file, flush=True)


class WriteSRT(SubtitlesWriter):
    extension: str = "srt"
    always_include_hours: bool = True
    decimal_marker: str = ","

    def write_result(self, result: dict, file: TextIO, options: Optional[dict] = None, **kwargs):
        for i, (start, end, text) in enumerate(self.iterate_result(result, options, **kwargs), start=1):
            print(f"{i}\n{start} --> {end}\n{text}\n", file=file, flush=True)


class WriteTSV(ResultWriter):
    """Tab-separated start / end (integer milliseconds) / text."""

    extension: str = "tsv"

    def write_result(self, result: dict, file: TextIO, options: Optional[dict] = None, **kwargs):
        print("start", "end", "text", sep="\t", file=file)
        for segment in result["segments"]:
            print(round(1000 * segment["start"]), file=file, end="\t")
            print(round(1000 * segment["end"]), file=file, end="\t")
            print(segment["text"].strip().replace("\t", " "), file=file, flush=True)


class WriteJSON(ResultWriter):
    extension: str = "json"

    def write_result(self, result: dict, file: TextIO, options: Optional[dict] = None, **kwargs):
        json.dump(result, file, ensure_ascii=False)


def get_writer(output_format: str, output_dir: str) -> Callable[[dict, TextIO, dict], None]:
    writers = {"txt": WriteTXT, "vtt": WriteVTT, "srt": WriteSRT, "tsv": WriteTSV, "json": WriteJSON}
    if output_format == "all":
        all_writers = [writer(output_dir) for writer in writers.values()]

        def write_all(result: dict, output_name: str, options: Optional[dict] = None, **kwargs):
            for writer in all_writers:
                writer(result, output_name, options, **kwargs)

        return write_all
    return writers[output_format](output_dir)
