"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, the host-only entry points work, and device entry points fail loudly (there
is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ai00_server_b200 import capi, runtime, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200rwkv.h")).read()
    # the debug-build section (#ifdef B200RWKV_DEBUG ... #endif) is not part of the product library
    product_hdr = re.sub(r"#ifdef B200RWKV_DEBUG.*?#endif", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b200rwkv_[a-z0-9_]+)\s*\(", product_hdr))
    declared -= {"b200rwkv_status", "b200rwkv_info", "b200rwkv_engine", "b200rwkv_options"}
    assert len(declared) >= 30
    lib = capi.lib()
    bound = {n for n, _, _ in capi.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert getattr(lib, name) is not None
    for name, _, _ in capi.DEBUG_SYMBOLS:            # and the product library really does not carry the debug entries
        assert not hasattr(lib, name)


def test_product_library_ignores_the_environment():
    """The bring-up switches (B200RWKV_*) exist only in the debug build: every getenv in the engine sources sits inside an
    `#ifdef B200RWKV_DEBUG` block (the CUDA runtime linked into the library reads its own CUDA_* variables)."""
    csrc = os.path.join(ROOT, "ai00_server_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if not f.endswith((".cu", ".cuh")):
            continue
        src = open(os.path.join(csrc, f)).read()
        outside = re.sub(r"#ifdef B200RWKV_DEBUG.*?#e(?:lse|ndif)", "", src, flags=re.S)
        assert "getenv(" not in outside, f


def test_info_from_st_host_only():
    for preset, ver in (("tiny5", 5), ("tiny6", 6), ("tiny7", 7)):
        info = capi.info_from_st(synth.make_st(preset, 0))
        s = synth.PRESETS[preset]
        assert info["version"] == ver
        assert (info["num_layer"], info["num_emb"], info["num_hidden"], info["num_vocab"]) == (s.L, s.C, s.F, s.V)
        assert (info["num_head"], info["head_size"]) == (s.H, 64)
    i6 = capi.info_from_st(synth.make_st("tiny6", 0))
    assert (i6["time_mix_adapter"], i6["time_decay_adapter"]) == (32, 64)


def test_malformed_st_is_an_error_not_a_crash():
    junk = np.frombuffer(b"\x10\x00\x00\x00\x00\x00\x00\x00{\"a\":1}        ", dtype=np.uint8).copy()
    with pytest.raises(capi.B200Error) as ei:
        capi.info_from_st(junk)
    assert ei.value.code == capi.ERR_INVALID
    with pytest.raises(capi.B200Error):
        capi.info_from_st(np.zeros(4, np.uint8))


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(capi.B200Error) as ei:
        runtime.Model(synth.make_st("tiny6", 0), max_batch=2, token_chunk_size=16)
    assert ei.value.code == capi.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_bad_precision_rejected():
    """precision 0 = f16 activations, 1 = f32-exact activations (web-rwkv Bundle::<f32>); anything else is invalid."""
    st = synth.make_st("tiny6", 0)
    h = C.c_void_p()
    code = capi.lib().b200rwkv_create(capi.ptr(st), st.size, 0, 2, 16, 7, C.byref(h))
    assert code == capi.ERR_INVALID
    assert not h.value


def _edit_header(st: np.ndarray, fn) -> np.ndarray:
    import json
    import struct
    raw = st.tobytes()
    hlen = struct.unpack("<Q", raw[:8])[0]
    hdr = json.loads(raw[8:8 + hlen])
    fn(hdr)
    h2 = json.dumps(hdr, separators=(",", ":")).encode()
    return np.frombuffer(struct.pack("<Q", len(h2)) + h2 + raw[8 + hlen:], dtype=np.uint8).copy()


def test_st_tensor_sizes_are_validated():
    """A tensor whose byte range does not equal dtype x shape, a wrong rank, or an absurd shape must be a clean
    B200RWKV_ERR_INVALID from the host-only parser (the loader trusted shape-derived sizes)."""
    st = synth.make_st("tiny6", 0)
    assert capi.info_from_st(st)["version"] == 6

    def short_tensor(h):
        b, e = h["blocks.0.att.key.weight"]["data_offsets"]
        h["blocks.0.att.key.weight"]["data_offsets"] = [b, e - 64]

    def huge_shape(h):
        h["emb.weight"]["shape"] = [1 << 40, 1 << 40]

    def wrong_rank(h):
        t = h["blocks.0.att.time_first"]
        t["shape"] = [int(np.prod(t["shape"]))]

    def bad_dtype(h):
        h["emb.weight"]["dtype"] = "Q7"

    def deep_metadata(h):
        node = cur = {}
        for _ in range(200):
            cur["x"] = {}
            cur = cur["x"]
        h["__metadata__"] = node

    for fn in (short_tensor, huge_shape, wrong_rank, bad_dtype, deep_metadata):
        with pytest.raises(capi.B200Error) as ei:
            capi.info_from_st(_edit_header(st, fn))
        assert ei.value.code == capi.ERR_INVALID, fn.__name__


def test_runtime_chunking_bookkeeping():
    """Runtime.infer mirrors web-rwkv: <= token_chunk_size tokens per call, Last rows only once a
    slot's run is exhausted (reference run.rs:1134-1155).  Engine calls are stubbed."""
    calls = []

    class Stub:
        info = {"num_vocab": 8}

        def infer_raw(self, slots, ntok, toks, opts):
            calls.append((list(slots), list(ntok), list(toks), list(opts)))
            return [np.zeros((nt if o == capi.OPTION_FULL else (1 if o == capi.OPTION_LAST else 0), 8), np.float32)
                    for nt, o in zip(ntok, opts)]

    rt = runtime.Runtime(Stub())
    inp = runtime.RnnInput([runtime.RnnInputBatch([1, 2, 3, 4, 5], runtime.RnnOption.Last),
                            runtime.RnnInputBatch([], runtime.RnnOption.Last),
                            runtime.RnnInputBatch([7, 8], runtime.RnnOption.Full)], 4)
    seen_rows = {0: 0, 2: 0}
    while inp.num_token() > 0:
        inp, out = rt.infer(inp)
        for b, o in enumerate(out):
            if not o.is_empty():
                seen_rows[b] += o.data.shape[0]
    assert seen_rows == {0: 1, 2: 2}
    assert calls[0] == ([0], [4], [1, 2, 3, 4], [capi.OPTION_NONE])
    assert calls[1] == ([0, 2], [1, 2], [5, 7, 8], [capi.OPTION_LAST, capi.OPTION_FULL])


def test_read_state_host_only():
    """`vN::read_state` (reference lib.rs:378-389): a state-tuned model / `.state` file -> the [L, N+2, C] state tensor; the
    oracle builds the same tensor from the same file (row 1+i, column h*N+j <- time_state[h][i][j])."""
    import dataclasses
    import json
    import struct
    from oracle import rwkv_numpy as O
    shp = dataclasses.replace(synth.PRESETS["tiny6"], time_state=True)
    st = synth.make_st(shp, 0)
    info = capi.info_from_st(st)
    got = runtime.read_state(info, st)
    want = O.Oracle(O.parse_st(st), "f16").state_init()
    assert got.shape == want.shape and np.array_equal(got, want) and np.abs(got[:, 1:65]).max() > 0
    assert np.all(got[:, 0] == 0) and np.all(got[:, 65] == 0)
    # a stand-alone `.state` file: only the time_state tensors, stored as F32
    w = O.parse_st(st)
    names = [n for n in w if n.endswith("att.time_state")]
    hdr, blobs, off = {"__metadata__": {"format": "pt"}}, [], 0
    for n in names:
        b = w[n].astype(np.float32).tobytes()
        hdr[n] = {"dtype": "F32", "shape": list(w[n].shape), "data_offsets": [off, off + len(b)]}
        blobs.append(b); off += len(b)
    h = json.dumps(hdr).encode()
    img = np.frombuffer(struct.pack("<Q", len(h)) + h + b"".join(blobs), np.uint8).copy()
    assert np.array_equal(runtime.read_state(info, img), want)
    with pytest.raises(capi.B200Error) as ei:           # a model without time_state is not a state file
        runtime.read_state(info, synth.make_st("tiny6", 0))
    assert ei.value.code == capi.ERR_INVALID


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirrors of the two ABI structs have the layout a C compiler gives include/b200rwkv.h (a plain C translation unit:
    the header must stay C, not C++)."""
    import ctypes as C
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200rwkv.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(b200rwkv_options), offsetof(b200rwkv_options, devices),\n'
                   '  offsetof(b200rwkv_options, lora_st), offsetof(b200rwkv_options, quant_layers), sizeof(b200rwkv_info)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    O = capi.Options
    assert got == [C.sizeof(O), O.devices.offset, O.lora_st.offset, O.quant_layers.offset, C.sizeof(capi.Info)]


def test_c_host_program_links_and_calls_the_library(tmp_path):
    """A plain C host (no Python, no torch) includes include/b200rwkv.h, links libb200rwkv.so and calls the host-only entry
    points the Rust shim would call first (`Loader::info`, then `create`, which must fail loudly on a box without a GPU)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "ai00_server_b200")
    st = synth.make_st("tiny7", 0)
    stp = tmp_path / "m.st"
    stp.write_bytes(st.tobytes())
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "b200rwkv.h"
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)malloc((size_t)n);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) return 3;
    b200rwkv_info info;
    int32_t rc = b200rwkv_info_from_st(buf, (size_t)n, &info);
    printf("%d %d %d %d %d %d %d %d\n", rc, info.version, info.num_layer, info.num_emb, info.num_hidden, info.num_vocab, info.num_head, info.head_size);
    b200rwkv_engine* e = NULL;
    rc = b200rwkv_create(buf, (size_t)n, 0, 2, 32, 0, &e);
    printf("%d %s\n", rc, b200rwkv_last_error(NULL));
    if (e) b200rwkv_destroy(e);
    return 0;
}
''')
    exe = tmp_path / "host"
    subprocess.run([gcc, "-std=c99", "-Wall", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), "-L", libdir, "-lb200rwkv",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe), str(stp)], check=True, capture_output=True, text=True).stdout.splitlines()
    s = synth.PRESETS["tiny7"]
    assert [int(x) for x in out[0].split()] == [0, 7, s.L, s.C, s.F, s.V, s.H, s.N]
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        rc, msg = out[1].split(" ", 1)
        assert int(rc) < 0 and "CUDA" in msg
