"""luxgi_b200 — B200-native DDGI probe-update engine (the SDF-traced branch of flwmxd/LuxGI's DDGI pass).

Layout: csrc/ (sm_100a kernels + C++ host behind the C ABI of include/luxddgi.h), ddgi.py (ctypes host binding that
mirrors the reference's systems), abi.py (POD mirrors), scenes.py (synthetic input fixtures), build.py (nvcc recipe).
"""
__version__ = "0.1.0"
