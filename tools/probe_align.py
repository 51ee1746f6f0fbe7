"""Sweeps of the tcgen05 probe (adec_probe_mma_ex): A-operand placement (non-aligned window rows) and several unordered issuer warps."""
import ctypes, sys
sys.path.insert(0, ".")
from audiodec_b200 import _lib
lib = _lib.load()
tf, ms = ctypes.c_double(), ctypes.c_double()
print("kind NT  off pitch step issuers  TFLOP/s   clk/MMA")
for nt in (32, 64, 128, 256):
    for off, pitch, step, ni in ((0, 128, 0, 1), (1, 186, 9, 1), (0, 128, 0, 2), (0, 128, 0, 3), (0, 128, 0, 4), (1, 186, 9, 3)):
        rc = lib.adec_probe_mma_ex(0, 1, nt, 6000, off, pitch, step, ni, ctypes.byref(tf), ctypes.byref(ms))
        clk = 128 * nt * 16 * 2 / (tf.value * 1e12 / 148 / 1.965e9) if rc == 0 else -1
        print(f"f16 {nt:4d} {off:4d} {pitch:5d} {step:4d} {ni:4d} {tf.value:12.1f} {clk:9.1f}")
