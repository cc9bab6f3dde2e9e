"""multithreadedgameengine_b200 — B200-native spatial + physics hot path for the WeedJS engine.

Only what the path needs lives here: ``csrc/`` (hand-written sm_100a CUDA kernels behind the
C ABI of include/weedgpu.h), the ctypes binding (``binding``), the host-side mirror of the
reference's worker/engine interface (``engine``, ``components``) and the synthetic scenes of
BASELINE.json (``scenes``).  There is no CPU fallback: without libweedgpu.so and a CUDA
device, engine construction raises.
"""
__version__ = "0.1.0"
