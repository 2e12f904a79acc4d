"""`transcribe_audio` tool module for the reference daemon's plugin API.

The reference's only plugin interface is the tool system: a module exports `TOOL`, a `Tool(spec, execute)`
whose `execute(**arguments) -> str` returns a JSON string (/root/reference/daemon/tools/base.py:23-105; tools are
registered lazily by module path, /root/reference/daemon/tools/registry.py:185-239; failures are reported as
`{"error": ..., "status": "error"}` JSON rather than raised, e.g. /root/reference/daemon/tools/ocr/ocr_document.py:186-204).
No transcription tool exists there; this module is the one a maintainer would register with

    registry.register_lazy("transcribe_audio", "whisper_mlx_b200.tool", "TOOL")

When the reference package is importable its own `tool` decorator builds `TOOL`; otherwise structurally identical
frozen dataclasses are used, so the object answers `.name`, `.description`, `.parameters`, `.to_schema()` and
`.execute(...)` either way.  The work itself is `whisper_mlx_b200.transcribe` (no CPU fallback: without the CUDA
library the call reports the error).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Callable, Optional

try:  # inside the reference tree: use its types so isinstance checks in its registry hold
    from daemon.tools.base import Tool, ToolSpec, tool  # type: ignore
except Exception:  # standalone: same fields and methods as daemon/tools/base.py:23-105

    @dataclass(frozen=True)
    class ToolSpec:
        name: str
        description: str
        parameters: dict

        def to_schema(self) -> dict:
            return {"name": self.name, "description": self.description, "parameters": self.parameters}

    @dataclass(frozen=True)
    class Tool:
        spec: ToolSpec
        execute: Callable[..., Any]

        @property
        def name(self) -> str:
            return self.spec.name

        @property
        def description(self) -> str:
            return self.spec.description

        @property
        def parameters(self) -> dict:
            return self.spec.parameters

        def to_schema(self) -> dict:
            return self.spec.to_schema()

    def tool(name: str, description: str, parameters: dict) -> Callable[[Callable[..., Any]], Tool]:
        def decorator(fn):
            return Tool(spec=ToolSpec(name=name, description=description, parameters=parameters), execute=fn)

        return decorator


DEFAULT_MODEL_ENV = "B200W_MODEL"
DEFAULT_MODEL = "mlx-community/whisper-large-v3-mlx"  # the model `./run` names (/root/reference/run:4)


def _error(msg: str) -> str:
    return json.dumps({"error": msg, "status": "error"})


@tool(
    name="transcribe_audio",
    description="""Transcribe a speech recording to text with Whisper on the local B200 GPU.

Accepts a 16-bit PCM WAV file (any other container needs ffmpeg on the host). Returns the full text, the
detected language and timestamped segments. Long recordings are processed in 30 s windows.""",
    parameters={
        "type": "object",
        "properties": {
            "file_path": {"type": "string", "description": "Path to the audio file to transcribe"},
            "language": {"type": "string", "description": "Language code (e.g. 'en'). Default: auto-detect"},
            "task": {"type": "string", "description": "'transcribe' (default) or 'translate' (to English)"},
            "model": {"type": "string",
                      "description": f"Model directory or HF repo. Default: ${DEFAULT_MODEL_ENV} or {DEFAULT_MODEL}"},
            "max_segments": {"type": "integer", "description": "Return at most this many segments (text is always complete)"},
        },
        "required": ["file_path"],
    },
)
def transcribe_audio(file_path: str, language: Optional[str] = None, task: str = "transcribe", model: Optional[str] = None,
                     max_segments: Optional[int] = None) -> str:
    path = Path(file_path).expanduser().resolve()
    if not path.exists():
        return _error(f"File not found: {file_path}")
    if task not in ("transcribe", "translate"):
        return _error(f"Unsupported task: {task}. Supported: transcribe, translate")
    try:
        from . import transcribe  # the package attribute of that name is the function

        result = transcribe(str(path), path_or_hf_repo=model or os.environ.get(DEFAULT_MODEL_ENV, DEFAULT_MODEL),
                            language=language, task=task, verbose=None,
                            # the flags `./run` passes (/root/reference/run:5-6)
                            condition_on_previous_text=False, hallucination_silence_threshold=1.0)
    except Exception as e:  # reported, not raised: the daemon feeds the string back to the model
        return _error(f"{type(e).__name__}: {e}")
    segments = [{"id": s["id"], "start": round(float(s["start"]), 2), "end": round(float(s["end"]), 2),
                 "text": s["text"].strip()} for s in result["segments"]]
    if max_segments is not None and max_segments >= 0:
        segments = segments[:max_segments]
    return json.dumps({"status": "success", "file": str(path), "language": result["language"],
                       "text": result["text"].strip(), "segment_count": len(result["segments"]), "segments": segments,
                       "char_count": len(result["text"].strip())})


TOOL = transcribe_audio
