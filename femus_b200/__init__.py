"""femus_b200: B200 (sm_100a) backend for the FEMuS assembly + geometric-multigrid hot path.

The product is ``libfemus_b200.so`` (hand-written CUDA kernels behind the C ABI of
``include/femus_b200.h``) plus the C++ host layer in ``femus_b200/host``; :mod:`femus_b200.capi`
is the ctypes harness used by tests and benchmarks."""
