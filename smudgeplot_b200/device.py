"""Layer A of the C ABI driven from Python: torch supplies device memory and streams (plumbing
only), every computation is one of our CUDA kernels in libhetmers_b200.so.

`DeviceTable` holds the structure-of-arrays table of DESIGN.md §3 in torch tensors:
    keys  int64[n]   (bit pattern of the left-aligned uint64 packed k-mer, bases 0..31)
    keys_lo int64[n] (bases 32..63, k > 32 only)
    cnt   int16[n]   (bit pattern of the uint16 count)
    deg   uint8[n+]  (the reference's `Pair` incidence array, PloidyPlot.c:163)
    bucket int32/int64[(1<<bits)+1]
    filter int32[2^fbits/32]  (prefix presence bitmap probed by pass 1)
    up    int32/int64[hi-lo]  (upper partner recorded by pass 1)
    plot  int64[1001*501]
"""
from __future__ import annotations

import torch

from . import _lib


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _stream():
    return torch.cuda.current_stream().cuda_stream


class DeviceTable:
    def __init__(self, kmer: int, keys: torch.Tensor, cnt: torch.Tensor, bits: int | None = None,
                 fbits: int | None = None, keys_lo: torch.Tensor | None = None, force_idx64: bool = False):
        assert keys.is_cuda and keys.dtype == torch.int64 and keys.is_contiguous()
        assert (kmer > 32) == (keys_lo is not None), "k > 32 needs the second key word (keys_lo)"
        self.keys_lo = keys_lo
        assert cnt.is_cuda and cnt.dtype == torch.int16 and cnt.is_contiguous()
        self.L = _lib.lib()
        self.kmer = kmer
        self.keys, self.cnt = keys, cnt
        self.n = keys.numel()
        self.device = keys.device
        self.idx64 = int(self.n >= 0xFFFFFFF0 or force_idx64)     # 64-bit offsets (tables >= 2^32 entries)
        self.idx_dtype = torch.int64 if self.idx64 else torch.int32
        self.bits = bits if bits is not None else self.L.hm_pick_bucket_bits(self.n)
        self.fbits = fbits if fbits is not None else self.L.hm_pick_filter_bits(self.n)
        self.bucket = None
        self.filter = None
        self.deg = None
        self.up = None
        self.lo = self.hi = 0
        self.launches = 0
        self.shards = None          # _lib.Shards when the incidence array is sharded over GPUs
        self.deg_ptr = None         # raw device pointer overriding self.deg (IPC-shared allocation)

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def from_records(cls, kmer: int, ibyte: int, records: torch.Tensor, index: torch.Tensor,
                     first: int = 0, n_total: int | None = None, out=None, out_lo=None):
        """Unpack raw FastK part records (uint8[n*pbyte], on the device) holding table ordinals
        [first, first+n) into SoA tensors; `index` is the stub index int64[1<<8*ibyte] on the
        device.  With `out=(keys, cnt)` (full-table tensors) the shard is written in place."""
        L = _lib.lib()
        kbyte = (kmer + 3) >> 2
        pbyte = kbyte - ibyte + 2
        n = records.numel() // pbyte
        lv = klo = None
        if out is None:
            keys = torch.empty(n, dtype=torch.int64, device=records.device)
            cnt = torch.empty(n, dtype=torch.int16, device=records.device)
            kv, cv = keys, cnt
            if kmer > 32:
                klo = lv = torch.empty(n, dtype=torch.int64, device=records.device)
        else:
            keys, cnt = out
            kv, cv = keys[first:first + n], cnt[first:first + n]
            if kmer > 32:
                lv = out_lo[first:first + n]
        with torch.cuda.device(records.device):
            _lib.check(L.hm_k_unpack_records(_ptr(records), n, first, _ptr(index), ibyte, kmer,
                                             _ptr(kv), _ptr(lv), _ptr(cv), _stream()))
        if out is None:
            t = cls(kmer, keys, cnt, keys_lo=klo)
            t.launches += 1
            return t
        return None

    def build_index(self):
        self.bucket = torch.empty((1 << self.bits) + 1, dtype=self.idx_dtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.hm_k_build_bucket_index(_ptr(self.keys), self.n, self.bits,
                                                      _ptr(self.bucket), self.idx64, _stream()))
        self.filter = torch.empty(self.L.hm_filter_words(self.fbits), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.hm_k_build_filter(_ptr(self.keys), self.n, self.fbits, _ptr(self.filter), _stream()))
        self.launches += 2
        return self

    # ---- the two passes -----------------------------------------------------------------
    def alloc_work(self, lo: int = 0, hi: int | None = None):
        hi = self.n if hi is None else hi
        self.lo, self.hi = lo, hi
        self.deg = torch.zeros((self.n + 4) & ~3, dtype=torch.uint8, device=self.device)
        self.up = torch.empty(max(hi - lo, 1), dtype=self.idx_dtype, device=self.device)
        self.plot = torch.zeros(_lib.PLOT_CELLS, dtype=torch.int64, device=self.device)
        return self

    def _deg(self):
        return self.deg_ptr if self.deg_ptr is not None else _ptr(self.deg)

    def _shards(self):
        import ctypes as C
        return C.byref(self.shards) if self.shards is not None else None

    def pass1(self):
        """neighbour search + degree (hm_k_pass1_degree) over [lo,hi); deg must be zero."""
        with torch.cuda.device(self.device):
            _lib.check(self.L.hm_k_pass1_degree(_ptr(self.keys), _ptr(self.keys_lo), _ptr(self.cnt), self.n, _ptr(self.bucket),
                                                self.bits, self.idx64, _ptr(self.filter), self.fbits, self.kmer,
                                                self.lo, self.hi,
                                                self._deg(), _ptr(self.up), self._shards(), _stream()))
        self.launches += 1

    def pass2(self):
        """isolated pairs -> plot (hm_k_pass2_plot) over [lo,hi); accumulates into self.plot."""
        with torch.cuda.device(self.device):
            _lib.check(self.L.hm_k_pass2_plot(_ptr(self.cnt), self._deg(), _ptr(self.up), self.idx64,
                                              self.lo, self.hi, _ptr(self.plot), self._shards(), _stream()))
        self.launches += 1

    def scan(self):
        """both passes on one GPU; -> plot int64[1001,501] (device tensor)"""
        if self.bucket is None:
            self.build_index()
        if self.deg is None:
            self.alloc_work()
        else:
            self.deg.zero_()
            self.plot.zero_()
        self.pass1()
        self.pass2()
        return self.plot.view(_lib.SMAX + 1, _lib.PLOT_W)

    # ---- examine_table pieces -----------------------------------------------------------
    def min_count(self, frst: int, last: int) -> int:
        out = torch.full((1,), 0x8000, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.hm_k_min_count(_ptr(self.cnt), frst, last, _ptr(out), _stream()))
        self.launches += 1
        return int(out.item())

    def find(self, queries: torch.Tensor, queries_lo: torch.Tensor | None = None) -> torch.Tensor:
        pos = torch.empty(queries.numel(), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.hm_k_find_keys(_ptr(self.keys), _ptr(self.keys_lo), self.n, _ptr(self.bucket), self.bits,
                                             self.idx64, _ptr(queries), _ptr(queries_lo), queries.numel(), _ptr(pos),
                                             _stream()))
        self.launches += 1
        return pos
