"""Result writers: host mirror of `mlx_whisper/writers.py` (UPSTREAM; `-f txt --output-name` in
/root/reference/run:3; restated in SURVEY.md A.7)."""
from __future__ import annotations

import json
import os
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
    always_include_hours: bool
    decimal_marker: str

    def iterate_result(self, result: dict, options: Optional[dict] = None, **kwargs):
        for segment in result["segments"]:
            start = self.format_timestamp(segment["start"])
            end = self.format_timestamp(segment["end"])
            text = segment["text"].strip().replace("-->", "->")
            yield start, end, text

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
