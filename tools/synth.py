"""Seeded synthetic FastK k-mer tables (test / bench input generator; not product code).

Genome model of SURVEY.md §8(d): G uniform-random bases; ploidy P haplotypes = copies of
haplotype 0 with iid SNPs at rate `het` per extra haplotype (alt base uniform over the other 3);
every k-mer of every haplotype on both strands; count(x) ~ Poisson(cov/P * occurrences(x)) drawn
from a counter-based RNG keyed on canonical(x), so count(x) == count(rc(x)); entries with
count < L are dropped, so the table is *trimmed* and *symmetric* exactly as the reference
requires (PloidyPlot.c:1167-1230) and neither implementation needs FastK's Logex/Symmex.

Everything is integer arithmetic on int64 torch tensors (splitmix64 mixing, integer Poisson
thresholds), so the same (k, G, P, het, cov, L, seed) gives the same table on CPU and on GPU.
Keys are LEFT-ALIGNED packed 2-bit k-mers held as the bit pattern of an int64 (base i in bits
63-2i..62-2i; a=0 c=1 g=2 t=3), returned sorted in *unsigned* order.
"""
from __future__ import annotations

import math

import numpy as np
import torch

_M64 = (1 << 64) - 1
_SIGN = -(1 << 63)


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """logical shift right on int64 bit patterns"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def mix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 bit patterns (wrapping arithmetic)."""
    x = x + _s64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _s64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _s64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _hash_idx(n: int, salt: int, device) -> torch.Tensor:
    i = torch.arange(n, dtype=torch.int64, device=device)
    return mix64(i ^ _s64(mix_int(salt)))


def mix_int(v: int) -> int:
    """splitmix64 on a python int"""
    v = (v + 0x9E3779B97F4A7C15) & _M64
    v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & _M64
    return v ^ (v >> 31)


def reverse2(x: torch.Tensor) -> torch.Tensor:
    """reverse the order of the 32 2-bit fields of int64 bit patterns"""
    x = ((x >> 2) & _s64(0x3333333333333333)) | ((x & _s64(0x3333333333333333)) << 2)
    x = ((x >> 4) & _s64(0x0F0F0F0F0F0F0F0F)) | ((x & _s64(0x0F0F0F0F0F0F0F0F)) << 4)
    x = ((x >> 8) & _s64(0x00FF00FF00FF00FF)) | ((x & _s64(0x00FF00FF00FF00FF)) << 8)
    x = ((x >> 16) & _s64(0x0000FFFF0000FFFF)) | ((x & _s64(0x0000FFFF0000FFFF)) << 16)
    x = ((x >> 32) & _s64(0x00000000FFFFFFFF)) | (x << 32)
    return x


def revcomp_left(x: torch.Tensor, k: int) -> torch.Tensor:
    """reverse complement of left-aligned packed k-mers (k <= 32), result left-aligned"""
    r = reverse2(~x)                       # right-aligned rc with the (complemented) pad on top
    if k < 32:
        r = (r & ((1 << (2 * k)) - 1)) << (64 - 2 * k)
    return r


def sort_unsigned(x: torch.Tensor):
    """ascending sort of int64 bit patterns in unsigned order"""
    s, _ = torch.sort(x ^ _SIGN)
    return s ^ _SIGN


def _poisson_tables(lam1: float, mmax: int, cmax: int):
    """flat integer threshold table: entry (m, c) = m<<53 | floor(CDF_{m*lam1}(c) * 2^53)"""
    flat = np.empty((mmax + 1, cmax), dtype=np.int64)
    for m in range(mmax + 1):
        lam = lam1 * m
        if m == 0:
            cdf = np.ones(cmax)
        else:
            c = np.arange(cmax, dtype=np.float64)
            logp = -lam + c * math.log(lam) - np.array([math.lgamma(v + 1.0) for v in c])
            cdf = np.minimum(np.cumsum(np.exp(logp)), 1.0)
        thr = np.floor(cdf * float(1 << 53)).astype(np.int64)
        thr = np.minimum(thr, (1 << 53) - 1)
        thr[-1] = (1 << 53) - 1            # the last bin absorbs the tail
        flat[m] = (np.int64(m) << 53) | thr
    return flat.reshape(-1)


def synth_table(k: int, G: int, ploidy: int = 2, het: float = 0.01, cov: float = 40.0,
                L: int = 4, seed: int = 1, device="cpu", key_range=None, extra_hom_repeats: int = 0):
    """-> (keys int64[n] left-aligned bit patterns in unsigned ascending order, cnt int32[n]).

    key_range=(lo, hi): keep only k-mers whose top-24-bit prefix is in [lo, hi) (used to generate
    one rank's shard of a large table; the union over a partition of [0, 2^24) is the full table).
    extra_hom_repeats: append that many tandem copies of the first 10k bases to every haplotype
    (gives entries with occurrences > ploidy, i.e. >2 neighbours per position in rare cases).
    """
    assert 1 <= k <= 64 and G >= k
    if k > 32:
        return _synth_table_long(k, G, ploidy, het, cov, L, seed, device, key_range, extra_hom_repeats)
    dev = torch.device(device)
    base0 = _hash_idx(G, seed * 1000003 + 1, dev) & 3
    if extra_hom_repeats:
        rep = base0[: min(10000, G)]
        base0 = torch.cat([base0] + [rep] * extra_hom_repeats)
        G = base0.numel()
    n_k = G - k + 1
    het_thr = int(het * float(1 << 53))
    chunks = []
    for h in range(ploidy):
        if h == 0:
            b = base0
        else:
            hsh = _hash_idx(G, seed * 1000003 + 17 * h + 5, dev)
            is_snp = _lsr(hsh, 11) < het_thr
            delta = 1 + (hsh & 0x7FF) % 3
            b = torch.where(is_snp, (base0 + delta) & 3, base0)
        v = torch.zeros(n_k, dtype=torch.int64, device=dev)
        for j in range(k):
            v = (v << 2) | b[j:j + n_k]
        if k < 32:
            v = v << (64 - 2 * k)
        for w in (v, revcomp_left(v, k)):          # a rank's shard: filter before concatenating (memory)
            if key_range is not None:
                pre = _lsr(w, 40)
                w = w[(pre >= key_range[0]) & (pre < key_range[1])]
            chunks.append(w)
        del v, w
    allk = torch.cat(chunks)
    del chunks
    keys, occ = torch.unique(allk ^ _SIGN, sorted=True, return_counts=True)
    keys = keys ^ _SIGN
    del allk
    # counts: integer inverse-CDF Poisson keyed on canonical(x)
    rc = revcomp_left(keys, k)
    canon = torch.where((keys ^ _SIGN) < (rc ^ _SIGN), keys, rc)
    u = _lsr(mix64(canon ^ _s64(mix_int(seed * 7919 + 3))), 11)            # 53-bit uniform
    lam1 = cov / ploidy
    mmax = 64
    cmax = int(lam1 * mmax + 12 * math.sqrt(lam1 * mmax) + 16)
    flat = torch.from_numpy(_poisson_tables(lam1, mmax, cmax)).to(dev)
    m = torch.clamp(occ, max=mmax)
    cnt = torch.clamp(torch.searchsorted(flat, (m << 53) | u, right=True) - m * cmax, max=cmax - 1)
    cnt = torch.where(occ > mmax, (occ.double() * lam1).round().long(), cnt)
    cnt = torch.clamp(cnt, max=32767)
    keep = cnt >= L
    return keys[keep].contiguous(), cnt[keep].to(torch.int32).contiguous()


def revcomp_long(hi: torch.Tensor, lo: torch.Tensor, k: int):
    """reverse complement of left-aligned packed k-mers with 32 < k <= 64 held as (hi, lo) words"""
    a, b = reverse2(~lo), reverse2(~hi)          # all 64 slots reversed: the words swap
    sh = 2 * (64 - k)                            # complemented pad now on top: shift it out
    if sh == 0:
        return a, b
    return (a << sh) | _lsr(b, 64 - sh), b << sh


def _haplotype_bases(G, ploidy, het, seed, dev, extra_hom_repeats):
    base0 = _hash_idx(G, seed * 1000003 + 1, dev) & 3
    if extra_hom_repeats:
        rep = base0[: min(10000, G)]
        base0 = torch.cat([base0] + [rep] * extra_hom_repeats)
        G = base0.numel()
    het_thr = int(het * float(1 << 53))
    haps = [base0]
    for h in range(1, ploidy):
        hsh = _hash_idx(G, seed * 1000003 + 17 * h + 5, dev)
        is_snp = _lsr(hsh, 11) < het_thr
        delta = 1 + (hsh & 0x7FF) % 3
        haps.append(torch.where(is_snp, (base0 + delta) & 3, base0))
    return haps, G


def _synth_table_long(k, G, ploidy, het, cov, L, seed, device, key_range, extra_hom_repeats):
    """k in 33..64: same model, keys returned as int64[n,2] = (bases 0..31, bases 32..k-1) words"""
    dev = torch.device(device)
    haps, G = _haplotype_bases(G, ploidy, het, seed, dev, extra_hom_repeats)
    n_k = G - k + 1
    his, los = [], []
    for b in haps:
        hi = torch.zeros(n_k, dtype=torch.int64, device=dev)
        lo = torch.zeros(n_k, dtype=torch.int64, device=dev)
        for j in range(32):
            hi = (hi << 2) | b[j:j + n_k]
        for j in range(32, k):
            lo = (lo << 2) | b[j:j + n_k]
        lo = lo << (2 * (64 - k))
        rhi, rlo = revcomp_long(hi, lo, k)
        his += [hi, rhi]
        los += [lo, rlo]
    hi, lo = torch.cat(his), torch.cat(los)
    del his, los
    if key_range is not None:
        pre = _lsr(hi, 40)
        m = (pre >= key_range[0]) & (pre < key_range[1])
        hi, lo = hi[m], lo[m]
    pairs = torch.stack([hi ^ _SIGN, lo ^ _SIGN], dim=1)
    del hi, lo
    uniq, occ = torch.unique(pairs, dim=0, sorted=True, return_counts=True)
    del pairs
    hi, lo = uniq[:, 0] ^ _SIGN, uniq[:, 1] ^ _SIGN
    rhi, rlo = revcomp_long(hi, lo, k)
    fwd_smaller = ((hi ^ _SIGN) < (rhi ^ _SIGN)) | ((hi == rhi) & ((lo ^ _SIGN) < (rlo ^ _SIGN)))
    chi, clo = torch.where(fwd_smaller, hi, rhi), torch.where(fwd_smaller, lo, rlo)
    u = _lsr(mix64(mix64(chi ^ _s64(mix_int(seed * 7919 + 3))) ^ clo), 11)
    lam1 = cov / ploidy
    mmax = 64
    cmax = int(lam1 * mmax + 12 * math.sqrt(lam1 * mmax) + 16)
    flat = torch.from_numpy(_poisson_tables(lam1, mmax, cmax)).to(dev)
    m = torch.clamp(occ, max=mmax)
    cnt = torch.clamp(torch.searchsorted(flat, (m << 53) | u, right=True) - m * cmax, max=cmax - 1)
    cnt = torch.where(occ > mmax, (occ.double() * lam1).round().long(), cnt)
    cnt = torch.clamp(cnt, max=32767)
    keep = cnt >= L
    keys = torch.stack([hi[keep], lo[keep]], dim=1).contiguous()
    return keys, cnt[keep].to(torch.int32).contiguous()


def keys_to_u64_numpy(keys: torch.Tensor) -> np.ndarray:
    return keys.cpu().numpy().view(np.uint64)


def calibrate_G(k: int, target_nels: int, ploidy: int, het: float, cov: float, L: int) -> int:
    """genome length giving ~target_nels table entries (both strands, distinct, count >= L)."""
    lam1 = cov / ploidy
    # P(a k-mer of haplotype h>0 is novel) = 1-(1-het)^k; novel k-mers have occurrence 1
    novel = 1.0 - (1.0 - het) ** k

    def p_ge(lam):
        return 1.0 - sum(math.exp(-lam + c * math.log(lam) - math.lgamma(c + 1)) for c in range(L))

    # shared (all haplotypes) k-mers: roughly occurrence = ploidy*(1-novel)+... ; coarse model
    per_base = 2.0 * ((1.0 - novel) * p_ge(lam1 * ploidy) + ploidy * novel * p_ge(lam1))
    return max(k + 1, int(target_nels / per_base))


def write_table(name: str, k: int, keys: torch.Tensor, cnt: torch.Tensor, ibyte: int = 3,
                nparts: int = 1, cut_on_buckets: bool = True):
    from smudgeplot_b200 import fastk
    return fastk.write_ktab(name, k, keys_to_u64_numpy(keys), cnt.cpu().numpy().astype(np.uint16),
                            ibyte=ibyte, nparts=nparts, cut_on_buckets=cut_on_buckets)
