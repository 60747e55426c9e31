"""Drop-in for the reference's top-level module `depthwise_conv2d_implicit_gemm`
(depthwise_conv2d_implicit_gemm.py:52-66): same class name, constructor and state_dict
keys, still an nn.Conv2d subclass.  Add this directory to PYTHONPATH in place of
`cutlass/examples/19_large_depthwise_conv2d_torch_extension` (models/SLaK.py:9-10).
"""
from slak_b200.dwconv import DepthWiseConv2dImplicitGEMM  # noqa: F401

__all__ = ["DepthWiseConv2dImplicitGEMM"]
