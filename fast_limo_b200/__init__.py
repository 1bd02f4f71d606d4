"""fast_LIMO's per-scan registration path on B200 (sm_100a) behind the reference's call surface.

    api        ctypes mirror of include/flimo.h (libflimo_cuda.so; no CPU path): Mapper, MappingConfig, FilterConfig
    localizer  the two fast_limo::Localizer callbacks in sequence over `api`
    config     the reference's YAML parameter files -> MappingConfig / LocalizerConfig
    dist       one process per GPU: scan sharding + per-pass exchange of the 96-double partials
    synth      seeded synthetic worlds, scans, IMU streams (tests and bench; the reference ships no data)

Nothing is imported eagerly: `import fast_limo_b200.api` loads the shared library and fails loudly if it is missing.
"""
