"""What one tcgen05.mma kind::f16 instruction costs on a B200 SM, by shape and A-operand source (debug build,
b200rwkv_debug_mma_rate): cycles per instruction over 4096 back-to-back issues, and what that means as weight bytes per cycle
for the two ways a skinny projection can be laid on the tensor core -- weights as the A operand (M = 128 weight rows, N = token
rows: today's kernels) or weights as the B operand (N = weight rows, M = token rows padded to 64 / 128)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai00_server_b200 import capi
L = capi.debug_lib()
n = 4096
def run(M, N, ts):
    cyc = (C.c_int64 * 2)()
    capi.check(L.b200rwkv_debug_mma_rate(0, M, N, ts, n, cyc))
    return cyc[0] / n, cyc[1] / n
print("  M    N  A from   issue cyc/MMA  retired cyc/MMA | weights = A: B/cyc | weights = B: B/cyc | MAC/cyc")
for M in (128, 64):
    for ts in (0, 1):
        for N in (16, 32, 64, 128, 256):
            i, r = run(M, N, ts)
            print(f"{M:4d} {N:4d}  {'tmem' if ts else 'smem'}   {i:10.1f}   {r:12.1f}    | {M * 32 / r:14.1f}    | {N * 32 / r:14.1f}    | {M * N * 16 / r:8.0f}")
