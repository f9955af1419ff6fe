"""TEST INFRASTRUCTURE: runs a subset of tests/test_kernel_emulation.py against a ThreadSanitizer build of the emulator
harness (argv[1] = path of the .so; the interpreter must have been started with LD_PRELOAD=libtsan.so).  A missing
__syncthreads / __syncwarp in a kernel shows up as a data race between the host threads that play the CUDA threads."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tests.test_kernel_emulation as T  # noqa: E402

vp = ctypes.c_void_p
L = ctypes.CDLL(sys.argv[1])
L.emu_schwarz.restype = ctypes.c_int
L.emu_schwarz.argtypes = [ctypes.c_int64, vp, vp, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
L.emu_neumann.restype = None
L.emu_neumann.argtypes = [ctypes.c_int64, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int]
L.emu_stokes.restype = None
L.emu_stokes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                         ctypes.c_double, ctypes.c_int]
L.emu_ns.restype = None
L.emu_ns.argtypes = L.emu_stokes.argtypes
L.emu_pressure_faces.restype = None
L.emu_pressure_faces.argtypes = [ctypes.c_int64, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int]
T.test_schwarz_kernels_on_the_emulator(L, "linear", 8, "colours")
T.test_schwarz_ilu_kernels_on_the_emulator(L, "linear", 5, "colours")
T.test_level_scheduled_rows_equal_the_one_warp_walk(L, "ssor", "linear", 8, (2, 2, 2))
T.test_level_scheduled_rows_equal_the_one_warp_walk(L, "ilu", "linear", 8, (2, 2, 2))
T.test_stokes_kernel_on_the_emulator(L, "cube_tet10", "quadratic", "linear", False)
T.test_navier_stokes_kernel_on_the_emulator(L, "box", "biquadratic", "linear", False)
T.test_boundary_pressure_kernel_on_the_emulator(L, "cube_tet10", "quadratic")
print("tsan-run-finished")
