"""tiebrush_b200 — B200-native (sm_100a CUDA) implementation of TieBrush's merge/collapse hot path and
TieCov's coverage / junction / bedgraph accumulation, behind a C ABI (include/tiebrush_b200.h).

Python here is plumbing: `api.Context` mirrors the reference's call sites over ctypes, `sam` packs
records into SoA windows, `synth` generates the synthetic cohorts used by bench.py."""
__version__ = "0.1"
