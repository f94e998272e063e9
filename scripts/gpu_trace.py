"""Per-phase timing of the whole-step kernel (B200RWKV_TRACE=1): where does a decode step go?"""
import ctypes as C, os, sys
import numpy as np
os.environ["B200RWKV_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai00_server_b200 import capi, runtime, synth

preset = os.environ.get("B200RWKV_BENCH_PRESET", "v6-7b")
B = int(os.environ.get("B200RWKV_BENCH_BATCH", "16"))
st = synth.make_st(preset, 0)
m = runtime.Model(st, max_batch=B, token_chunk_size=64)
slots = list(range(B))
rng = np.random.default_rng(0)
for i in range(6):
    m.infer_raw(slots, [1] * B, rng.integers(1, 60000, B).tolist(), [0] * B)
buf = np.zeros(4 * 4096 * 12, np.uint64); types = np.zeros(4096, np.int32); n = C.c_int32(0)
capi.check(capi.lib().b200rwkv_debug_trace(m._h, capi.ptr(buf), buf.size, capi.ptr(types), C.byref(n)), m._h)
n = n.value
tr = buf[:4 * n * 12].reshape(4, n, 12).astype(np.int64)
names = {0: "EMBED", 1: "LN", 2: "GEMM", 3: "WKV", 4: "LNOUT", 5: "SMALLN", 6: "SMALLK"}
for c, label in enumerate(["cta0", "cta15", "cta74", "ctaLast"]):
    t = tr[c]
    start = np.concatenate([[t[0, 0]], t[:-1, 1]])       # phase start = previous barrier exit
    work = t[:, 0] - start
    wait = t[:, 1] - t[:, 0]
    tot = (t[-1, 0] - t[0, 0]) / 1e3
    print(f"{label}: total {tot:.1f} us")
    for ty in sorted(set(types[:n])):
        sel = types[:n] == ty
        print(f"   {names[ty]:6s} n={sel.sum():4d} work sum {work[sel].sum()/1e3:8.1f} us (avg {work[sel].mean()/1e3:6.2f})  barrier sum {wait[sel].sum()/1e3:8.1f} us (avg {wait[sel].mean()/1e3:6.2f})")
# layer 1 detail for cta0
t = tr[0]
start = np.concatenate([[t[0, 0]], t[:-1, 1]])
k0 = 1 + 9   # skip embed + layer 0 (9 phases)
for i in range(k0, k0 + 10):
    st = t[i, 4:12]
    extra = " ".join(f"{(x - start[i])/1e3:.2f}" if x > 0 else "-" for x in st)
    print(f"   ph{i} {names[types[i]]:6s} work {(t[i,0]-start[i])/1e3:7.2f} us  barrier {(t[i,1]-t[i,0])/1e3:7.2f} us   stamps(us from phase start): {extra}")
m.close()
