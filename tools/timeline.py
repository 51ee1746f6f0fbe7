"""Per-role timeline of one tensor-core conv CTA (library built with -DADEC_TIMELINE); full-GPU-sized problems so the
persistent kernel reaches steady state (tile #2 of CTA 1 is recorded)."""
import ctypes, sys, numpy as np
lib = ctypes.CDLL("audiodec_b200/lib/libaudiodec_b200_tl.so")
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
lib.adec_test_residual_unit.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
lib.adec_last_error.restype = ctypes.c_char_p; lib.adec_last_error.argtypes = [ctypes.c_void_p]
for C, d, T, B in ((32, 9, 48000, 8), (64, 9, 16000, 12), (128, 9, 4000, 24)):
    x = np.random.randn(B, C, T).astype(np.float32); w1 = (np.random.randn(C, C, 7) / 15).astype(np.float32); w2 = (np.random.randn(C, C, 1) / 8).astype(np.float32)
    st = np.zeros((B, C, 6 * d), np.float32); y = np.zeros((B, C, T), np.float32)
    print("=== RU C=%d d=%d T=%d B=%d" % (C, d, T, B), flush=True)
    rc = lib.adec_test_residual_unit(0, p(x), B, C, T, p(w1), p(w2), 7, d, p(st), p(y))
    print("rc", rc, lib.adec_last_error(None), flush=True)
