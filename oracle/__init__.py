"""CPU oracle for the Whisper transcription hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain NumPy / PyTorch-CPU fp32 restatement of the algorithm that
`mlx_whisper.transcribe` runs (reference call site: /root/reference/run:3-6).  The
implementation the reference reaches is the third-party PyPI package `mlx-whisper`
(source ml-explore/mlx-examples, whisper/mlx_whisper/, version undeclared and
unpinned in /root/reference/requirements.txt:1-41); it is not vendored under
/root/reference and cannot be installed here, so every function below restates the
published algorithm and is cross-pinned against the
independent `transformers` Whisper implementation that *is* installed in this image
(tests/test_oracle_vs_hf.py).

PARITY STATUS: "parity unpinned" against the reference itself -- the reference holds
no golden vector, known-answer test or fixture for this path
and its implementation is not runnable here.  The oracle is pinned by a second
source (HF transformers) and by committed golden vectors generated from it
(tests/golden/, generator tools/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (whisper-mlx_b200/) never does.
"""
