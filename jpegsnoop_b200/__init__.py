"""jpegsnoop_b200 — B200-native implementation of JPEGsnoop's scan-decode hot path
(CimgDecode::DecodeScanImg and callees) behind the reference's own CimgDecode API.

Layout:
  csrc/jsgpu_kernels.cu, jsgpu_api.cu   hand-written sm_100a kernels + the C-ABI (include/jsgpu.h)
  csrc/host/                            host C++: class CimgDecode (reference public surface),
                                        marker walk, flat C shim (include/jsimg.h)
  host.py                               ctypes faces: CimgDecode (1 image), BatchDecoder (n images)
  synth.py                              seeded synthetic baseline-JPEG generator (bench/test input)
  build.py                              in-tree nvcc build of libjsgpu.so
"""
from .host import CimgDecode, BatchDecoder, JsgpuError, parse_jpeg  # noqa: F401
