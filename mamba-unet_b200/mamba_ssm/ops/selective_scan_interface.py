"""Same dotted path as /root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py (the drop-in boundary).

Exports the three names Mamba-UNet's code touches (:14-152 of the reference file).  The fused
MambaInnerFn* / BiMambaInnerFn variants (:155-633) belong to the 1-D Mamba block, which Mamba-UNet never
instantiates, and are out of scope.
"""
from selscan_b200.ops import SelectiveScanFn, selective_scan_fn, selective_scan_ref  # noqa: F401

__all__ = ["SelectiveScanFn", "selective_scan_fn", "selective_scan_ref"]
