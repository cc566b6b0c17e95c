"""Scale probe (run under gpurun): build a synthetic index on the device, time the stages, run device-resident
count/locate batches, verify them with the size-independent checks and measure the gather32 roofline."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sview_fmindex_b200 import EncodingTable, FmIndex, FmIndexBuilder, IndexType, _ffi, synth  # noqa: E402

L = _ffi.lib()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
batches = [int(float(x)) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1e6", "1e7", "1e8"])]
plen = int(sys.argv[3]) if len(sys.argv) > 3 else 20


def chk(rc):
    if rc:
        raise RuntimeError(f"rc={rc} {L.svfm_last_error()}")


torch.cuda.init()
t0 = time.time()
d_text = torch.empty(n, dtype=torch.uint8, device="cuda")
alpha = np.frombuffer(synth.NUCLEOTIDES, dtype=np.uint8)
chk(L.svfm_bench_synth_text(d_text.data_ptr(), n, 42, alpha.ctypes.data, 4, 0, 0, None))
torch.cuda.synchronize()
print(f"synth text {n}: {time.time()-t0:.2f}s", flush=True)
enc = EncodingTable.from_symbols([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
it = IndexType(32, 3, 64, True)
b = FmIndexBuilder(n, 5, enc, it)
b.kmer_size, b.sampling_ratio = 3, 2
size = b.blob_size()
d_blob = torch.empty(size, dtype=torch.uint8, device="cuda")
t0 = time.time()
b.build_device(d_text.data_ptr(), d_blob.data_ptr(), size)
torch.cuda.synchronize()
print(f"build_device: {time.time()-t0:.2f}s blob={size}", flush=True)
t0 = time.time()
ix = FmIndex.load_device(d_blob.data_ptr(), size, it)
print(f"load_device: {time.time()-t0:.2f}s peak_mem={torch.cuda.max_memory_allocated()/1e9:.1f}GB", flush=True)
del d_blob
torch.cuda.empty_cache()
info = ix.info()
print("text_len", info.text_len, "sentinel", info.sentinel_index)

# gather roofline
for mb in (687, 2688):
    buf = torch.empty(mb * 10**6, dtype=torch.uint8, device="cuda")
    sps, ms = C.c_double(), C.c_double()
    chk(L.svfm_bench_gather32(buf.data_ptr(), buf.numel(), 1 << 28, 3, 1, C.byref(sps), C.byref(ms), None))
    print(f"gather32 over {mb} MB: {sps.value/1e9:.2f} Gsectors/s = {sps.value*32/1e12:.2f} TB/s ({ms.value:.2f} ms)", flush=True)
    del buf

sess = C.c_void_p()
chk(L.svfm_session_create(ix.handle, C.byref(sess)))
chk(L.svfm_session_set_timing(sess, 1))
for B in batches:
    d_pats = torch.empty(B * plen, dtype=torch.uint8, device="cuda")
    d_starts = torch.empty(B, dtype=torch.int64, device="cuda")
    chk(L.svfm_bench_synth_patterns(d_text.data_ptr(), n, d_pats.data_ptr(), d_starts.data_ptr(), B, plen, 42 + B, None))
    torch.cuda.synchronize()
    d_counts = torch.empty(B, dtype=torch.int32, device="cuda")
    d_offs = torch.empty(B + 1, dtype=torch.int64, device="cuda")
    for rep in range(3):
        t0 = time.time()
        chk(L.svfm_count_batch_device(sess, d_pats.data_ptr(), None, B, plen, 0, d_counts.data_ptr()))
        chk(L.svfm_session_sync(sess))
        tc = time.time() - t0
        t0 = time.time()
        dpos, total = C.c_void_p(), C.c_uint64()
        chk(L.svfm_locate_batch_device(sess, d_pats.data_ptr(), None, B, plen, 0, d_offs.data_ptr(), C.byref(dpos), C.byref(total)))
        chk(L.svfm_session_sync(sess))
        tl = time.time() - t0
        ms = (C.c_double * 8)()
        ln = (C.c_uint64 * 8)()
        chk(L.svfm_session_get_timing(sess, ms, ln, 1))
        print(f"B={B} rep{rep}: count {tc*1e3:.2f} ms ({B/tc/1e6:.1f} Mpat/s)  locate {tl*1e3:.2f} ms ({B/tl/1e6:.1f} Mpat/s) total={total.value} "
              f"phases ms={[round(x,2) for x in ms[:6]]}", flush=True)
    viol = (C.c_uint64 * 3)()
    dig = C.c_uint64()
    chk(L.svfm_bench_verify_locate(d_text.data_ptr(), n, d_pats.data_ptr(), plen, B, d_starts.data_ptr(), d_offs.data_ptr(),
                                   dpos, 32, enc.table.ctypes.data, viol, C.byref(dig), None))
    s_, d_ = C.c_uint64(), C.c_uint64()
    chk(L.svfm_bench_count_digest(d_counts.data_ptr(), 32, B, C.byref(s_), C.byref(d_), None))
    print(f"  verify: violations={list(viol)} digest={dig.value} count_sum={s_.value} (total {total.value})", flush=True)
    del d_pats, d_starts, d_counts, d_offs
L.svfm_session_destroy(sess)
