"""hephaestus_b200: B200-native (sm_100a) H(curl) curl-curl + mass operator apply and
device-resident PCG behind Hephaestus's plug points.

Everything numerical runs in libhcurl_b200.so (hand-written CUDA, C ABI in
include/hcurl_b200.h).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import HcurlError  # noqa: F401
