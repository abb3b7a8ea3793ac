"""B200-native charge/light readout chain of larnd-sim (hot path only).

Drop-in modules with the reference's call surface (``kernel[grid, block](*arrays)``):
``quenching, drifting, pixels_from_track, detsim, fee, lightLUT, light_sim`` -- see INTEGRATION.md.
All arithmetic runs in ``csrc/liblarndsim_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/larndsim_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
__all__ = ["consts", "quenching", "drifting", "pixels_from_track", "detsim", "fee", "lightLUT", "light_sim",
           "chain", "rng", "synth"]
