"""solat_optics_b200 -- B200 (sm_100a) implementation of the solat-optics beam-simulation hot path.

Drop-in modules (same names and call signatures as the reference package `solat_optics`):
    ot_geo, ray_trace_int, latrt_geo, solat_apert_field, far_field_latrt, phase_correct
plus `maps` (whole-map plans, multi-GPU sharding) which replaces the reference's MPI fan-out.
All ray tracing and field integration runs in libsolat_b200.so (hand-written CUDA, C ABI in
include/solat_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
