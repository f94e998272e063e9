"""Is an L2 prefetch issued while HBM is idle still in L2 when the next launch streams the same bytes?  (b200rwkv_debug_prefetch)
Rows: streamed MB, consumer CTAs, prefetched blocks per consumer (32 KB each), mode, idle us -> us of the streaming launch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai00_server_b200 import capi
L = capi.debug_lib()        # python -m ai00_server_b200.build --debug
def run(mb, consumers, pf_grid, skip, nblk, mode, idle_us, reps=8):
    ms = (C.c_float * 2)()
    capi.check(L.b200rwkv_debug_prefetch(0, mb, consumers, pf_grid, skip, nblk, mode, idle_us, reps, ms))
    return ms[0] * 1e3, ms[1] * 1e3
for mb, cons in ((135.3, 128), (33.6, 128), (151.0, 144)):
    per = int(mb * 1e6 / cons / 32768)
    print(f"--- {mb} MB over {cons} CTAs ({per} blocks each)")
    base = run(mb, cons, 128, 0, 0, 0, 15.0)
    print(f"  no prefetch, idle 15us:                 stream {base[0]:7.2f} us   (prefetch+idle+stream {base[1]:7.2f})")
    for nblk in (4, 8, 16, 24, 32):
        if nblk > per: continue
        for mode in (0, 1, 2):
            for idle in (15.0,):
                t = run(mb, cons, 128, 0, nblk, mode, idle)
                print(f"  nblk {nblk:2d} ({nblk * cons * 32768 / 1e6:6.1f} MB) mode {mode} idle {idle:4.1f}us: stream {t[0]:7.2f} us   (total {t[1]:7.2f})")
    t = run(mb, cons, 128, 5, min(16, per - 5), 0, 15.0)
    print(f"  nblk 16 after the first 5 (ring) blocks, mode 0:  stream {t[0]:7.2f} us")
    t = run(mb, cons, 1024, 0, per, 0, 8.0)
    print(f"  everything, issued by 1024 CTAs, idle 8us:        stream {t[0]:7.2f} us")
    t = run(mb, cons, 128, 0, min(16, per), 0, 0.0)
    print(f"  nblk 16, no idle time (back to back):             stream {t[0]:7.2f} us (total {t[1]:7.2f})")
