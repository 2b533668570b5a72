"""pgslam_b200 — B200-native scan-registration hot path behind libpointmatcher's
plugin surface, as Ellon/pgslam binds to it (types.h:19-27).

Importing this package does not load CUDA; `pgslam_b200.pm` dlopens the in-tree
libpgslam_b200.so on first use and raises if it (or a GPU) is missing — there is
no CPU fallback by design.
"""
__all__ = ["pm", "synth", "build", "dist"]
__version__ = "0.1.0"
