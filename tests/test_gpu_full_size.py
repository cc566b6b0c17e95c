"""Full-size parity of the BASELINE.json configurations on one B200.

The oracle cannot search 10^7..10^8 patterns in seconds, so every configuration is checked three ways:
  * size-independent properties of the WHOLE batch, computed on the device (include/svfm_bench.h):
    every reported position really matches its pattern symbol by symbol after encoding, every pattern's own source
    position is in its list, no list of a pattern cut from the text is empty, count_i == out_offs[i+1]-out_offs[i];
  * bit-exact comparison with the CPU oracle (counts, CSR offsets, positions in SA-row order) on the first `sample`
    patterns, against the very blob the GPU searched (copied back to the host);
  * the host-buffer entry point on a slice of the batch must reproduce the device-resident result.
The index is built on the GPU (svfm_build_device: byte-identical to the oracle builder, tests/test_gpu_builder.py)."""
import ctypes as C
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DNA5 = [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]          # bench/src/build/mod.rs:30
DNA4 = [b"Aa", b"Cc", b"Gg", b"Tt"]                  # bench/src/build/sview_memory.rs:22-24 (T doubles as the wildcard)
AMINO = [bytes([c]) for c in b"ACDEFGHIKLMNPQRSTVWY"]


def _chk(L, rc):
    assert rc == 0, (rc, L.svfm_last_error())


def _device_copy(dst_ptr: int, src_ptr, nbytes: int):
    """cudaMemcpy device-to-device from a raw pointer handed out by the C ABI (cuda-python)."""
    if nbytes == 0:
        return
    from cuda.bindings import runtime as cudart
    src = src_ptr.value if hasattr(src_ptr, "value") else int(src_ptr)
    err, = cudart.cudaMemcpy(dst_ptr, src, nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert int(err) == 0, err


def _run_config(oracle, *, n, alphabet, rare, symbols, wildcard, pos_bits, planes, vec_bits, k, r, plen, batch, sample, seed=42):
    import torch

    import sview_fmindex_b200 as fm
    from sview_fmindex_b200 import _ffi
    po = oracle
    L = _ffi.lib()
    free_b, _ = torch.cuda.mem_get_info()
    need = n * (40 if n > 2**31 else 30)
    if free_b < need:
        pytest.skip(f"needs ~{need >> 30} GiB of device memory for the suffix sort, {free_b >> 30} GiB free")
    enc = fm.EncodingTable.from_symbols_with_wildcard(symbols) if wildcard else fm.EncodingTable.from_symbols(symbols)
    it = fm.IndexType(pos_bits, planes, vec_bits, True)
    d_text = torch.empty(n, dtype=torch.uint8, device="cuda")
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    _chk(L, L.svfm_bench_synth_text(d_text.data_ptr(), n, seed, alpha.ctypes.data, len(alphabet), rare, ord("X"), None))
    b = fm.FmIndexBuilder(n, enc.symbol_count(), enc, it)
    b.kmer_size, b.sampling_ratio = k, r
    size = b.blob_size()
    d_blob = torch.empty(size, dtype=torch.uint8, device="cuda")
    t0 = time.time()
    b.build_device(d_text.data_ptr(), d_blob.data_ptr(), size)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    ix = fm.FmIndex.load_device(d_blob.data_ptr(), size, it)
    host_blob = po.aligned_empty(size)
    torch.from_numpy(host_blob).copy_(d_blob)
    del d_blob
    torch.cuda.empty_cache()
    assert ix.info().text_len == n

    d_pats = torch.empty(batch * plen, dtype=torch.uint8, device="cuda")
    d_starts = torch.empty(batch, dtype=torch.int64, device="cuda")
    _chk(L, L.svfm_bench_synth_patterns(d_text.data_ptr(), n, d_pats.data_ptr(), d_starts.data_ptr(), batch, plen, seed + 7, None))
    sess = C.c_void_p()
    _chk(L, L.svfm_session_create(ix.handle, C.byref(sess)))
    np_pos = np.uint32 if pos_bits == 32 else np.uint64
    t_pos = torch.int32 if pos_bits == 32 else torch.int64
    d_counts = torch.empty(batch, dtype=t_pos, device="cuda")
    d_offs = torch.empty(batch + 1, dtype=torch.int64, device="cuda")
    dpos, total = C.c_void_p(), C.c_uint64()
    times = []
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        _chk(L, L.svfm_count_batch_device(sess, d_pats.data_ptr(), None, batch, plen, 0, d_counts.data_ptr()))
        _chk(L, L.svfm_session_sync(sess))
        t1 = time.time()
        _chk(L, L.svfm_locate_batch_device(sess, d_pats.data_ptr(), None, batch, plen, 0, d_offs.data_ptr(), C.byref(dpos), C.byref(total)))
        _chk(L, L.svfm_session_sync(sess))
        times.append((t1 - t0, time.time() - t1))
    # ---- whole batch: size-independent properties ----------------------------------------------------------------
    viol = (C.c_uint64 * 3)()
    dig = C.c_uint64()
    _chk(L, L.svfm_bench_verify_locate(d_text.data_ptr(), n, d_pats.data_ptr(), plen, batch, d_starts.data_ptr(), d_offs.data_ptr(),
                                       dpos, pos_bits, enc.table.ctypes.data, viol, C.byref(dig), None))
    assert list(viol) == [0, 0, 0], list(viol)
    offs = d_offs.cpu().numpy().astype(np.uint64)
    counts = d_counts.cpu().numpy().view(np_pos).astype(np.uint64)
    assert int(offs[-1]) == total.value and offs[0] == 0
    assert np.array_equal(np.diff(offs), counts)          # count == number of located positions, pattern by pattern
    assert counts.min() >= 1                               # every pattern was cut from the text
    # ---- oracle on a sample, same blob -----------------------------------------------------------------------------
    ora = po.OracleFmIndex.load(host_blob, po.IndexType(pos_bits, planes, vec_bits, True))
    pats = d_pats[:sample * plen].cpu().numpy().reshape(sample, plen)
    oc, oo, op_, _ = ora.locate_batch(pats, threads=8)
    assert np.array_equal(counts[:sample], oc)
    assert np.array_equal(offs[:sample + 1], oo)
    n_pos = int(oo[-1])
    pos_t = torch.empty(n_pos, dtype=t_pos, device="cuda")
    _device_copy(pos_t.data_ptr(), dpos, n_pos * (pos_bits // 8))
    assert np.array_equal(pos_t.cpu().numpy().view(np_pos).astype(np.uint64), op_.astype(np.uint64))   # SA-row order
    # ---- host-buffer entry point on a slice ------------------------------------------------------------------------
    m = min(batch, 3_000_000)
    host_pats = d_pats[:m * plen].cpu().numpy().reshape(m, plen)
    h_offs, h_pos = ix.locate_batch(host_pats)
    assert np.array_equal(h_offs, offs[:m + 1])
    assert np.array_equal(ix.count_batch(host_pats).astype(np.uint64), counts[:m])
    pos_m = torch.empty(int(offs[m]), dtype=t_pos, device="cuda")
    _device_copy(pos_m.data_ptr(), dpos, int(offs[m]) * (pos_bits // 8))
    assert np.array_equal(h_pos, pos_m.cpu().numpy().view(np_pos))
    L.svfm_session_destroy(sess)
    tc, tl = times[-1]
    print(f"\n[full-size] n={n:.3g} {it} S={enc.symbol_count()} k={k} r={r}: build {t_build:.2f} s, {batch} x {plen}: "
          f"count {tc * 1e3:.1f} ms ({batch / tc / 1e6:.0f} M/s), locate {tl * 1e3:.1f} ms ({batch / tl / 1e6:.0f} M/s), "
          f"{total.value} occurrences")
    ix.close()


def test_cfg1_cfg2_dna_block3(oracle):
    """configs[0]/[1]: 1 Gbp, u32, Block3<u64>, S=5, SA ratio 2, kLTS 3, 20 bp patterns."""
    _run_config(oracle, n=10**9, alphabet=b"ACGT", rare=0, symbols=DNA5, wildcard=False, pos_bits=32, planes=3, vec_bits=64,
                k=3, r=2, plen=20, batch=20_000_000, sample=20_000)


@pytest.mark.parametrize("ratio", [2, 16])
def test_cfg3_block2_reads(oracle, ratio):
    """configs[2]: 1 Gbp Block2<u64> ACGT-only index, 150 bp reads, SA sampling ratio 2 vs 16 (LF-walk length sweep)."""
    _run_config(oracle, n=10**9, alphabet=b"ACGT", rare=0, symbols=DNA4, wildcard=False, pos_bits=32, planes=2, vec_bits=64,
                k=3, r=ratio, plen=150, batch=2_000_000, sample=5_000)


@pytest.mark.parametrize("planes", [5, 6])
def test_cfg4_protein(oracle, planes):
    """configs[3]: 500 Maa protein text (+0.1 % X -> wildcard), 20 symbols + wildcard, Block5<u64> and the widest
    Block6<u64>, 12-mers, kLTS 3."""
    _run_config(oracle, n=5 * 10**8, alphabet=b"ACDEFGHIKLMNPQRSTVWY", rare=1000, symbols=AMINO, wildcard=True, pos_bits=32,
                planes=planes, vec_bits=64, k=3, r=2, plen=12, batch=20_000_000, sample=20_000)


def test_cfg5_human_scale_u64(oracle):
    """configs[4]: 3.1 Gbp, u64 positions, Block3<u64>, S=5, SA ratio 2, kLTS 3, 32 bp patterns."""
    _run_config(oracle, n=31 * 10**8, alphabet=b"ACGT", rare=0, symbols=DNA5, wildcard=False, pos_bits=64, planes=3, vec_bits=64,
                k=3, r=2, plen=32, batch=20_000_000, sample=10_000)
