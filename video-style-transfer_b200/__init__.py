"""B200-native ReCoNet / RTNSTV frame path.

Sub-packages mirror the reference's module surface:
  vst_b200.reconet.network / .utilities   <- RC/network.py, RC/utilities.py
  vst_b200.rtnstv.network / .vgg19 / .utilities <- RT/network.py, RT/vgg19.py, RT/utilities.py
Everything computes through the C-ABI library `csrc/libvst_b200.so` (include/vst_b200.h);
there is no CPU fallback.
"""
__version__ = "0.1.0"
