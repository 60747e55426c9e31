"""slak_b200: B200-native (sm_100a) implementation of SLaK's large-kernel depthwise
convolution hot path behind the reference's operator API.  See DESIGN.md."""
__version__ = "0.1.0"
