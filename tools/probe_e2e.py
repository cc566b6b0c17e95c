"""End-to-end probe (run under gpurun): raw pinned H2D/D2H bandwidth, then svfm_locate_batch / svfm_count_batch on
pinned host buffers for a few chunk sizes / worker counts.  SVFM_TRACE=1 prints per-chunk host timestamps."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sview_fmindex_b200 import EncodingTable, FmIndex, FmIndexBuilder, IndexType, _ffi, synth  # noqa: E402

L = _ffi.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
B = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**8
chunks = [int(float(x)) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["16777216"])]
plen = 20


def chk(rc):
    if rc:
        raise RuntimeError(f"rc={rc} {L.svfm_last_error()}")


torch.cuda.init()
# raw copy bandwidth
hp = torch.empty(2 * 10**9, dtype=torch.uint8).pin_memory()
dp = torch.empty(2 * 10**9, dtype=torch.uint8, device="cuda")
hq = torch.empty(12 * 10**8, dtype=torch.uint8).pin_memory()
dq = torch.empty(12 * 10**8, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time(); dp.copy_(hp, non_blocking=True); torch.cuda.synchronize(); t1 = time.time()
    hq.copy_(dq, non_blocking=True); torch.cuda.synchronize(); t2 = time.time()
    with torch.cuda.stream(s1):
        dp.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2):
        hq.copy_(dq, non_blocking=True)
    torch.cuda.synchronize(); t3 = time.time()
    print(f"H2D 2GB {2/(t1-t0):.1f} GB/s ({(t1-t0)*1e3:.1f} ms)  D2H 1.2GB {1.2/(t2-t1):.1f} GB/s ({(t2-t1)*1e3:.1f} ms)  both {(t3-t2)*1e3:.1f} ms", flush=True)
del hp, dp, hq, dq

d_text = torch.empty(n, dtype=torch.uint8, device="cuda")
alpha = np.frombuffer(synth.NUCLEOTIDES, dtype=np.uint8)
chk(L.svfm_bench_synth_text(d_text.data_ptr(), n, 42, alpha.ctypes.data, 4, 0, 0, None))
enc = EncodingTable.from_symbols([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
it = IndexType(32, 3, 64, True)
b = FmIndexBuilder(n, 5, enc, it)
b.kmer_size, b.sampling_ratio = 3, 2
size = b.blob_size()
d_blob = torch.empty(size, dtype=torch.uint8, device="cuda")
b.build_device(d_text.data_ptr(), d_blob.data_ptr(), size)
ix = FmIndex.load_device(d_blob.data_ptr(), size, it)
del d_blob
d_pats = torch.empty(B * plen, dtype=torch.uint8, device="cuda")
chk(L.svfm_bench_synth_patterns(d_text.data_ptr(), n, d_pats.data_ptr(), None, B, plen, 4242, None))
torch.cuda.synchronize()
p = L.svfm_host_alloc(B * plen)
arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(B * plen,))
torch.from_numpy(arr).copy_(d_pats)
cap = B + B // 4 + 1024
h_offs_p = L.svfm_host_alloc((B + 1) * 8)
h_pos_p = L.svfm_host_alloc(cap * 4)
h_cnt_p = L.svfm_host_alloc(B * 4)
total = C.c_uint64()
for ch in chunks:
    chk(L.svfm_set_tuning(_ffi.SVFM_TUNE_CHUNK, ch))
    for rep in range(4):
        t0 = time.time()
        chk(L.svfm_locate_batch(ix.handle, p, None, B, plen, 0, h_offs_p, h_pos_p, cap, C.byref(total)))
        t1 = time.time()
        chk(L.svfm_count_batch(ix.handle, p, None, B, plen, 0, h_cnt_p))
        t2 = time.time()
        print(f"chunk={ch} rep{rep}: locate_batch {(t1-t0)*1e3:.1f} ms ({B/(t1-t0)/1e6:.0f} Mpat/s)  count_batch {(t2-t1)*1e3:.1f} ms "
              f"({B/(t2-t1)/1e6:.0f} Mpat/s) total={total.value}", flush=True)
