"""Drop-in import path for the reference's `mamba_ssm` package, restricted to the hot path Mamba-UNet uses.

Put `<repo>/mamba-unet_b200` on PYTHONPATH (ahead of any installed mamba_ssm) and the reference's
`from mamba_ssm.ops.selective_scan_interface import selective_scan_fn, selective_scan_ref`
(code/networks/mamba_sys.py:17-20) resolves to the sm_100a kernels; no causal_conv1d / transformers needed.
"""
__version__ = "1.0.1+b200"
