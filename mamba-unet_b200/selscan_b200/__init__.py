"""B200-native (sm_100a) selective scan for Mamba-UNet's VSS blocks: host side of libselscan_b200.so."""
from .ops import SelectiveScanFn, selective_scan_fn, selective_scan_ref  # noqa: F401

__all__ = ["SelectiveScanFn", "selective_scan_fn", "selective_scan_ref"]
