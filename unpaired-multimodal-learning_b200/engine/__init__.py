"""Host-side mirror of the reference's ``vision_language/engine`` package, restricted to the
hot path: heads, optimizer/scheduler factories, hyper-parameter presets, feature-bank datasets and
loaders, and the config parser.  Names and argument meaning follow the reference so that callers of
``engine.*`` can switch over; the arithmetic runs in libuml_b200 CUDA kernels."""
