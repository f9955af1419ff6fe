"""TEST INFRASTRUCTURE: the multigrid orchestration on the emulator (tests/test_kernel_emulation.py) against a
ThreadSanitizer build of the emulated library (argv[1]); start the interpreter with LD_PRELOAD=libtsan.so."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tests.test_kernel_emulation as T  # noqa: E402

L = ctypes.CDLL(sys.argv[1])
L.emu_mg_run.restype = ctypes.c_int
L.emu_stokes_plan.restype = ctypes.c_int
L.b2_last_error.restype = ctypes.c_char_p
for case in sys.argv[2:]:
    if case == "stokes_plan":
        T.test_stokes_plan_entry_points_on_the_emulator(L, 1)
    else:
        T.test_multigrid_orchestration_on_the_emulator(L, case)
print("tsan-run-finished")
