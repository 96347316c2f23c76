"""CPU oracle for the selective-scan hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this package.  The product (mamba-unet_b200/) never does: it fails loudly when its CUDA library is missing.
"""
from .oracle import (  # noqa: F401
    build_oracle,
    oracle_fwd,
    oracle_bwd,
    oracle_threads,
    make_inputs,
)
