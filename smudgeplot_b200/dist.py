"""One-process-per-GPU sharding of the hetmers scan (torch.distributed = plumbing; NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests of the host logic).

Path partition: the sorted table is cut into `world` contiguous
shards on prefix-bucket boundaries ("canonical-prefix buckets" of BASELINE.json's north_star).
Rank r loads / owns shard r, every rank ends with a full replica (shards broadcast over NVLink),
and rank r scans index range [lo_r, hi_r):

    pass 1 (own range, lower pair member does the book-keeping)  -> partial incidence array
    all-reduce(sum, uint8[n])                                    <- the one real exchange step
    pass 2 (own range)                                           -> partial plot
    all-reduce(sum, int64[1001*501])                             <- the reference's serial
                                                                    plot reduction, :1569-1575
The reference has no multi-process code at all; its only "collective" is that final sum.
"""
from __future__ import annotations

import time

import torch
import torch.distributed as dist


def prefix_partition(world: int, bits: int = 24):
    """equal split of the `bits`-bit prefix space: [(lo_r, hi_r)] for r in range(world)"""
    top = 1 << bits
    return [((top * r) // world, (top * (r + 1)) // world) for r in range(world)]


def shard_offsets(local_n: int, group=None, device=None):
    """all ranks' shard sizes -> (sizes list, offsets list, total)"""
    world = dist.get_world_size(group)
    t = torch.zeros(world, dtype=torch.int64, device=device)
    t[dist.get_rank(group)] = local_n
    dist.all_reduce(t, group=group)
    sizes = [int(x) for x in t.tolist()]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    return sizes, offs[:-1], offs[-1]


def _exchange_to_even_chunks(buf: torch.Tensor, sizes, offs, chunk: int, total: int, group=None):
    """buf (padded to world*chunk elements) holds this rank's shard [offs[r], offs[r]+sizes[r]) in place; afterwards
    it also holds everything of the even chunk [r*chunk, (r+1)*chunk): the slivers that belong to that chunk but
    were loaded by other ranks arrive by point-to-point copies (shards of a prefix partition are nearly even, so
    the slivers are small)"""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ops, keep = [], []
    for q in range(world):
        oq0, oq1 = offs[q], offs[q] + sizes[q]
        for r in range(world):
            if q == r:
                continue
            a, b = max(oq0, r * chunk), min(oq1, min((r + 1) * chunk, total))
            if a >= b or rank not in (q, r):
                continue
            peer = r if rank == q else q
            peer = dist.get_global_rank(group, peer) if group is not None else peer
            t = buf[a:b]
            keep.append(t)
            ops.append(dist.P2POp(dist.isend if rank == q else dist.irecv, t, peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def gather_table(local_keys: torch.Tensor, local_cnt: torch.Tensor, group=None, out=None,
                 local_lo: torch.Tensor | None = None, out_lo: torch.Tensor | None = None):
    """Assemble the full sorted table on every rank from per-rank shards (shard r = rank r's
    slice of the key space, so concatenation in rank order is the sorted table): the shards are evened
    out by small point-to-point copies and then ONE all-gather per array moves everything (round 1 did
    3 x world sequential broadcasts: 45 ms of a 95 ms end-to-end step at 8 GPUs).
    -> (keys_full, cnt_full, lo, hi) with [lo,hi) this rank's index range; with `local_lo` (second
    key word, k > 32) -> (keys_full, cnt_full, lo, hi, keys_lo_full).
    `out` buffers are used in place when they have room for world*ceil(total/world) elements."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes, offs, total = shard_offsets(local_keys.numel(), group, local_keys.device)
    chunk = (total + world - 1) // world if world > 0 else total
    padded = max(chunk * world, total)

    def room(t, dtype, device):
        if t is not None and t.numel() >= padded:
            return t
        return torch.empty(padded, dtype=dtype, device=device)

    kbuf = room(out[0] if out is not None else None, local_keys.dtype, local_keys.device)
    cbuf = room(out[1] if out is not None else None, local_cnt.dtype, local_cnt.device)
    lbuf = room(out_lo, local_lo.dtype, local_lo.device) if local_lo is not None else None
    lo, hi = offs[rank], offs[rank] + sizes[rank]
    pairs = [(kbuf, local_keys), (cbuf, local_cnt)] + ([(lbuf, local_lo)] if lbuf is not None else [])
    for buf, loc in pairs:
        if loc.numel() and loc.data_ptr() != buf[lo:hi].data_ptr():
            buf[lo:hi].copy_(loc)
    if world > 1 and total > 0:
        for buf, _ in pairs:
            raw = buf.view(torch.uint8)                     # counts travel as raw bytes (gloo has no int16 collectives)
            w = buf.element_size()
            _exchange_to_even_chunks(raw, [s_ * w for s_ in sizes], [o * w for o in offs], chunk * w, total * w, group)
            own = raw[rank * chunk * w:(rank + 1) * chunk * w].clone()
            try:
                dist.all_gather_into_tensor(raw[:world * chunk * w], own, group=group)
            except (RuntimeError, NotImplementedError):
                parts = [torch.empty_like(own) for _ in range(world)]
                dist.all_gather(parts, own, group=group)
                for r, pt in enumerate(parts):
                    raw[r * chunk * w:(r + 1) * chunk * w].copy_(pt)
    keys, cnt = kbuf[:total], cbuf[:total]
    if out is not None:                                      # caller's buffers too small for the padding: copy back
        if out[0].data_ptr() != kbuf.data_ptr():
            out[0][:total].copy_(keys)
            keys = out[0][:total]
        if out[1].data_ptr() != cbuf.data_ptr():
            out[1][:total].copy_(cnt)
            cnt = out[1][:total]
    if lbuf is not None:
        klo = lbuf[:total]
        if out_lo is not None and out_lo.data_ptr() != lbuf.data_ptr():
            out_lo[:total].copy_(klo)
            klo = out_lo[:total]
        return keys, cnt, lo, hi, klo
    return keys, cnt, lo, hi


def _cuda_view(ptr: int, nbytes: int, device) -> torch.Tensor:
    """uint8 tensor over a raw device allocation (no ownership) via __cuda_array_interface__"""
    class _Raw:
        pass
    r = _Raw()
    r.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(r, device=device)


def allreduce_deg(deg: torch.Tensor, group=None):
    """sum of the partial incidence arrays (uint8; no wrap: <= 3k <= 192 neighbours for k <= 64)"""
    dist.all_reduce(deg, op=dist.ReduceOp.SUM, group=group)
    return deg


def allreduce_plot(plot: torch.Tensor, group=None):
    dist.all_reduce(plot, op=dist.ReduceOp.SUM, group=group)
    return plot


def balanced_offsets(keys_full: torch.Tensor, world: int, depth: int = 3, per_candidate: float = 0.03):
    """Work ranges [off[r], off[r+1]) with equal pass-1 WORK rather than equal entry counts.
    Only neighbours y > x are probed, so an entry with base b at position p has 3-b candidates
    there ('a' three, 't' none) -- each ~3 % of an entry's cost.  For the first `depth` bases that
    number is the same for whole stretches of the sorted table, so shards cut by count are
    systematically uneven (measured at 8 GPUs: 9.8 vs 8.3 ms with equal counts; 9.4 vs 8.7 ms when
    only the first base is weighted).  The table is split into the 4^depth prefix classes, each
    weighted by its candidate count, and the cuts are placed on the cumulative weight."""
    n = keys_full.numel()
    if world <= 1:
        return [0, n]
    sign = -(1 << 63)
    ncls = 4 ** depth
    flipped = keys_full ^ sign                                   # unsigned order as signed
    marks = []
    for c in range(1, ncls):
        v = c << (64 - 2 * depth)                                # first key of prefix class c (unsigned)
        v = v - (1 << 64) if v >= (1 << 63) else v               # as the int64 bit pattern
        marks.append(v ^ sign)
    pos = torch.searchsorted(flipped, torch.tensor(marks, dtype=torch.int64, device=keys_full.device))
    bnd = [0] + [int(v) for v in pos.tolist()] + [n]
    w = []
    for c in range(ncls):
        cand = sum(3 - ((c >> (2 * (depth - 1 - p))) & 3) for p in range(depth))
        w.append(1.0 + per_candidate * cand)
    total = sum(w[c] * (bnd[c + 1] - bnd[c]) for c in range(ncls))
    offs, acc, c, at = [0], 0.0, 0, 0
    for r in range(1, world):
        target = total * r / world
        while c < ncls and acc + w[c] * (bnd[c + 1] - at) < target:
            acc += w[c] * (bnd[c + 1] - at)
            c += 1
            at = bnd[c] if c < ncls else n
        if c >= ncls:
            offs.append(n)
            continue
        step = int((target - acc) / w[c])
        acc += w[c] * step
        at += step
        offs.append(min(max(at, offs[-1]), n))
    return offs + [n]


def run_aligned_cuts(keys_full: torch.Tensor, kmer: int, world: int):
    """Shard cuts of the strand-symmetric scan: equal entry counts, each cut moved to the next RUN
    boundary (a run = entries sharing their first k/2 bases; all partners the run scan looks for lie in
    one run, so no pair straddles two shards).  Pure torch: every rank computes the same cuts from its
    replica.  -> [0, c1, ..., n]"""
    n = keys_full.numel()
    sh = 64 - 2 * (kmer >> 1)

    def pre(t):                                   # first k/2 bases as a non-negative number
        return (t >> sh) & ((1 << (64 - sh)) - 1) if sh > 0 else t

    cuts = [0]
    for r in range(1, world):
        c = max((n * r) // world, cuts[-1])
        if 0 < c < n:
            p0 = pre(keys_full[c - 1:c])
            while c < n:
                w = keys_full[c:c + 65536]
                d = torch.nonzero(pre(w) != p0)
                if d.numel():
                    c += int(d[0])
                    break
                c += w.numel()
        cuts.append(min(c, n))
    return cuts + [n]


def fingerprint_verdict(acc: torch.Tensor, group=None) -> bool:
    """acc = this rank's int64[4] fingerprint sums (device.DeviceTable.fingerprint over the entries it
    loaded, same seeds on every rank): symmetric iff the job-wide sums (mod 2^64) of {(x,cnt)} and
    {(rc x,cnt)} agree"""
    world = dist.get_world_size(group)
    parts = [torch.zeros_like(acc) for _ in range(world)]
    dist.all_gather(parts, acc, group=group)
    tot = [0, 0, 0, 0]
    for p in parts:
        for i, v in enumerate(p.tolist()):
            tot[i] = (tot[i] + v) & 0xFFFFFFFFFFFFFFFF
    return tot[0] == tot[2] and tot[1] == tot[3]


def common_seeds(device, group=None):
    """rank 0's fingerprint seeds for everybody"""
    import ctypes as C
    from . import _lib
    sd = (C.c_uint64 * 2)()
    _lib.lib().hm_symm_seeds(sd)
    t = torch.tensor([sd[0] - (1 << 64) if sd[0] >= (1 << 63) else sd[0],
                      sd[1] - (1 << 64) if sd[1] >= (1 << 63) else sd[1]], dtype=torch.int64, device=device)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    return [int(v) & 0xFFFFFFFFFFFFFFFF for v in t.tolist()]


def exchange_segments(seg: torch.Tensor, rank: int, group=None):
    """seg[world, m]: row `rank` is this rank's Bloom segment; fill in everybody else's (all-gather)"""
    own = seg[rank].clone()
    try:
        dist.all_gather_into_tensor(seg.view(-1), own, group=group)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(own) for _ in range(seg.shape[0])]
        dist.all_gather(parts, own, group=group)
        for r, p in enumerate(parts):
            seg[r].copy_(p)
    return seg


class PeerDeg:
    """The sharded incidence array of DESIGN.md §6 for a one-process-per-GPU job: every rank
    cudaMallocs its own full-length array through the C ABI (hm_dev_alloc), exports a CUDA IPC
    handle, and maps everybody else's (hm_ipc_open).  With it the exchange between the passes is
    fused into the kernels (remote atomics / remote loads over NVLink) and no collective moves the
    array.  `PeerDeg.create` returns None when IPC or native NVLink atomics are unavailable, and
    the caller falls back to the dense all-reduce."""

    def __init__(self, own, peers, nbytes, rank):
        self.own, self.peers, self.nbytes, self.rank = own, peers, nbytes, rank

    @classmethod
    def create(cls, n, device, group=None):
        import ctypes as C
        import os
        from . import _lib
        if os.environ.get("HETMERS_DENSE_EXCHANGE"):
            return None
        L = _lib.lib()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        nbytes = (n + 256) & ~255               # one buffer; the allocation holds TWO (see scan_on)
        ok = 1
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        with torch.cuda.device(device):
            if L.hm_dev_alloc(2 * nbytes, C.byref(own)) != 0 or L.hm_ipc_export(own, handle) != 0:
                ok = 0
        info = [None] * world
        dist.all_gather_object(info, (ok, handle.raw, torch.cuda.current_device() if device.index is None else device.index),
                               group=group)
        peers = [None] * world
        if all(i[0] for i in info):
            for r, (_, h, dev_r) in enumerate(info):
                if r == rank:
                    peers[r] = own.value
                    continue
                if not L.hm_p2p_native_atomics(device.index or 0, dev_r):
                    ok = 0
                    break
                ptr = C.c_void_p()
                with torch.cuda.device(device):
                    if L.hm_ipc_open(h, C.byref(ptr)) != 0:
                        ok = 0
                        break
                peers[r] = ptr.value
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            with torch.cuda.device(device):
                for r, pp in enumerate(peers):
                    if pp is not None and r != rank:
                        L.hm_ipc_close(C.c_void_p(pp))
                if own.value:
                    L.hm_dev_free(own)
            return None
        return cls(own.value, peers, nbytes, rank)

    def close(self, device):
        """unmap the peers' arrays and free the own one (call on every rank, after a barrier)"""
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        with torch.cuda.device(device):
            for r, pp in enumerate(self.peers):
                if pp is not None and r != self.rank:
                    L.hm_ipc_close(C.c_void_p(pp))
            if self.own:
                L.hm_dev_free(C.c_void_p(self.own))
        self.peers, self.own = [], None

    def shards(self, offsets, parity):
        """hm_shards over buffer `parity` (0/1) of every rank's allocation"""
        from . import _lib
        sh = _lib.Shards()
        sh.n_shards, sh.self_ = len(self.peers), self.rank
        for r, o in enumerate(offsets):
            sh.off[r] = o
        for r, pp in enumerate(self.peers):
            sh.deg[r] = pp + parity * self.nbytes
        return sh


class ShardedScan:
    """device-resident replica + this rank's work range; `scan()` = T_scan of SURVEY.md §8d"""

    def __init__(self, kmer, keys_full, cnt_full, lo, hi, group=None, keys_lo_full=None, path="auto"):
        import os
        from .device import DeviceTable
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.table = DeviceTable(kmer, keys_full, cnt_full, keys_lo=keys_lo_full).build_index(direct=False)
        self.load_lo, self.load_hi = lo, hi       # the shard this rank LOADED (equal prefix ranges)
        self.kmer = kmer
        self.n_total = keys_full.numel()
        self.bits = self.table.bits
        self.peer = None
        # is the whole table strand-symmetric?  every rank fingerprints the shard it loaded
        path = os.environ.get("HETMERS_PATH", path)
        self.seeds = common_seeds(keys_full.device, group)
        self.symmetric = kmer >= 2 and fingerprint_verdict(self.table.fingerprint(lo, hi, self.seeds), group)
        self.table.symmetric = self.symmetric
        self.path = "symm" if (self.symmetric and path != "direct") else "direct"
        if self.path == "symm":
            # strand-symmetric scan: run-aligned shards of equal size, Bloom segments all-gathered
            self.offsets = run_aligned_cuts(keys_full, kmer, self.world)
            lo, hi = self.offsets[self.rank], self.offsets[self.rank + 1]
            self.table.alloc_symm(lo, hi, self.table.make_symm_shards(self.offsets, self.rank) if self.world > 1 else None)
            self.lo, self.hi = lo, hi
            self.exchange = "all-gather of Bloom segments (NCCL) between run scan and resolve" if self.world > 1 else "none"
        else:
            self._init_direct(keys_full)
        self._barrier_t = torch.zeros(1, dtype=torch.int32, device=keys_full.device)
        self._step = 0
        if keys_full.is_cuda:
            self._side = torch.cuda.Stream(device=keys_full.device)
        if self.peer is not None:       # tensors over the two halves of the IPC allocation (no ownership)
            self._peer_views = [_cuda_view(self.peer.own + h * self.peer.nbytes, self.peer.nbytes, keys_full.device)
                                for h in (0, 1)]

    def _init_direct(self, keys_full):
        """direct passes (any table): work-balanced shards, incidence array sharded by owner"""
        self.table.build_filter()
        # the shard this rank SCANS and owns the incidence bytes of: equal work, not equal counts
        self.offsets = balanced_offsets(keys_full, self.world)
        lo, hi = self.offsets[self.rank], self.offsets[self.rank + 1]
        self.table.alloc_work(lo, hi)
        self.lo, self.hi = lo, hi
        self.peer = PeerDeg.create(self.n_total, keys_full.device, self.group) if self.world > 1 else None
        self.exchange = "peer-memory (remote atomics/loads over NVLink, CUDA IPC)" if self.peer else \
                        ("all-reduce(uint8[n]) via NCCL" if self.world > 1 else "none")

    @classmethod
    def from_synthetic(cls, k, G, ploidy, het, cov, L, seed, device, group=None):
        """every rank generates only its prefix-range shard of the seeded table, then the shards
        are exchanged (same flow as loading 1/world of the part files per rank)."""
        from tools import synth
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        rng = prefix_partition(world)[rank]
        keys, cnt = synth.synth_table(k, G, ploidy, het, cov, L, seed, device=device, key_range=rng)
        cnt16 = cnt.to(torch.int16)
        del cnt
        if k > 32:                               # two key words per k-mer
            khi, klo = keys[:, 0].contiguous(), keys[:, 1].contiguous()
            kf, cf, lo, hi, lf = gather_table(khi, cnt16, group, local_lo=klo)
            return cls(k, kf, cf, lo, hi, group, keys_lo_full=lf)
        kf, cf, lo, hi = gather_table(keys, cnt16, group)
        del keys, cnt16
        return cls(k, kf, cf, lo, hi, group)

    def close(self):
        if self.peer is not None:
            torch.cuda.synchronize()
            dist.barrier(self.group)              # nobody is still reading a peer's array
            self.peer.close(self.table.device)
            self.peer = None

    def scan(self, events=None):
        return self.scan_on(self.table, events)

    def scan_on(self, t, events=None):
        """one T_scan on table replica `t` (self.table for the resident scans, a freshly loaded
        replica in the e2e leg).  Peer mode keeps the phases ordered with two collectives per scan:
        a 4-byte all-reduce after pass 1 (every remote atomic has landed before any pass 2 reads),
        and the plot all-reduce after pass 2.  The incidence array is double-buffered: scan s uses
        buffer s&1 and clears the other one after the first barrier -- by then every rank has left
        pass 2 of scan s-1 (it could not have passed that scan's plot all-reduce otherwise), and
        nobody writes it before the plot all-reduce of scan s, which the clearing rank joins only
        after its memset (stream order)."""
        if self.path == "symm":
            return self._scan_symm(t, events)
        if self.peer is None:
            t.deg.zero_()
            t.plot.zero_()
            if events is not None:
                events[0].record()
            t.pass1()
            if events is not None:
                events[1].record()
            if self.world > 1:
                allreduce_deg(t.deg, self.group)
            t.pass2()
            if self.world > 1:
                allreduce_plot(t.plot, self.group)
            return t.plot
        par = self._step & 1
        self._step += 1
        t.deg_ptr = self.peer.own + par * self.peer.nbytes
        t.shards = self.peer.shards(self.offsets, par)
        if getattr(t, "_p2_scratch", None) is None:     # pass 2 batches its foreign look-ups through this
            t._p2_scratch = torch.empty(t.L.hm_pass2_scratch_bytes(t.hi - t.lo, t.idx64), dtype=torch.uint8,
                                        device=t.device)
        t.shards.scratch = t._p2_scratch.data_ptr()
        t.shards.scratch_bytes = t._p2_scratch.numel()
        ph = self._phase_events() if self.profile_phases else None
        t.plot.zero_()
        if events is not None:
            events[0].record()
        if ph: ph[0].record()
        t.pass1()
        if events is not None:
            events[1].record()
        if ph: ph[1].record()
        dist.all_reduce(self._barrier_t, group=self.group)            # all pass 1 kernels have landed
        if ph: ph[2].record()
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):                           # the next scan's buffer is cleared
            self._peer_views[par ^ 1].zero_()                         #   next to pass 2 ...
        t.pass2()
        if ph: ph[3].record()
        main.wait_stream(self._side)
        if ph: ph[4].record()
        allreduce_plot(t.plot, self.group)                            # ... and before anybody can use it
        if ph: ph[5].record()
        return t.plot

    def _scan_symm(self, t, events=None):
        """strand-symmetric scan, sharded: run scan of the own (run-aligned) range -> all-gather of the
        Bloom segments -> resolve of the own candidates -> plot all-reduce.  No peer memory needed."""
        ph = self._phase_events() if self.profile_phases else None
        t.plot.zero_()
        if events is not None:
            events[0].record()
        if ph: ph[0].record()
        t.runscan(mid_event=events[1] if events is not None else None)   # [0],[1]: the dominant kernel alone
        if ph: ph[1].record()
        if self.world > 1:
            exchange_segments(t.bloom_view(), self.rank, self.group)
        if ph: ph[2].record()
        t.resolve()
        if ph: ph[3].record(); ph[4].record()
        if self.world > 1:
            allreduce_plot(t.plot, self.group)
        if ph: ph[5].record()
        return t.plot

    def symm_ok(self) -> bool:
        """status words of the last symmetric scan on every rank (synchronises): all clean?"""
        bad = torch.zeros(1, dtype=torch.int32, device=self.table.device)
        if self.path == "symm":
            _, st = self.table.symm_status()
            bad[0] = int(st != 0)
        if self.world > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
        return int(bad.item()) == 0

    profile_phases = False

    def _phase_events(self):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        self._phases = getattr(self, "_phases", [])
        self._phases.append(ev)
        return ev

    def phase_ms(self):
        """mean ms of (pass1, barrier, pass2, zero-next, plot all-reduce) over the profiled scans"""
        torch.cuda.synchronize()
        rows = [[a.elapsed_time(b) for a, b in zip(ev[:-1], ev[1:])] for ev in getattr(self, "_phases", [])]
        if not rows:
            return None
        names = ["pass1", "barrier_after_pass1", "pass2", "zero_next_buffer", "plot_allreduce"] if self.path != "symm" else \
                ["runscan", "bloom_allgather", "resolve", "-", "plot_allreduce"]
        return {n: sum(r[i] for r in rows) / len(rows) for i, n in enumerate(names)}

    # ---- end-to-end from pinned host buffers (bench.py e2e leg) -----------------------------
    def measure_e2e(self, steps, warmup):
        """per step: H2D of this rank's shard of raw FastK records + the stub index, unpack,
        shard exchange, bucket index, scan, plot D2H on rank 0.  -> e2e dict (max over ranks)."""
        from . import _lib
        from .device import DeviceTable
        t = self.table
        dev = t.device
        k, n, lo, hi = self.kmer, self.n_total, self.load_lo, self.load_hi
        wlo, whi = self.lo, self.hi
        kbyte, ibyte = (k + 3) // 4, 3
        pbyte = kbyte - ibyte + 2
        m = hi - lo
        keys, cnt = t.keys, t.cnt
        rec = torch.empty((m, pbyte), dtype=torch.uint8, device=dev)
        ks, cs = keys[lo:hi], cnt[lo:hi].to(torch.int32) & 0xFFFF
        for j in range(ibyte, kbyte):
            rec[:, j - ibyte] = ((ks >> (56 - 8 * j)) & 0xFF).to(torch.uint8)
        rec[:, pbyte - 2] = (cs & 0xFF).to(torch.uint8)
        rec[:, pbyte - 1] = ((cs >> 8) & 0xFF).to(torch.uint8)
        index = torch.cumsum(torch.bincount((keys >> 40) & 0xFFFFFF, minlength=1 << 24), 0)
        h_rec = torch.empty(rec.numel(), dtype=torch.uint8, pin_memory=True)
        h_rec.copy_(rec.view(-1))
        h_idx = torch.empty(1 << 24, dtype=torch.int64, pin_memory=True)
        h_idx.copy_(index)
        h_plot = torch.empty(_lib.PLOT_CELLS, dtype=torch.int64, pin_memory=True)
        del rec, index, ks, cs
        # the replica built by the timed call reuses no state of self.table
        d_rec = torch.empty(h_rec.numel(), dtype=torch.uint8, device=dev)
        d_idx = torch.empty(1 << 24, dtype=torch.int64, device=dev)
        k2 = torch.empty(n + self.world, dtype=torch.int64, device=dev)     # room for the all-gather's even chunks
        c2 = torch.empty(n + self.world, dtype=torch.int16, device=dev)

        def call():
            d_rec.copy_(h_rec, non_blocking=True)
            d_idx.copy_(h_idx, non_blocking=True)
            DeviceTable.from_records(k, ibyte, d_rec, d_idx, first=lo, out=(k2, c2))
            gather_table(k2[lo:hi], c2[lo:hi], self.group, out=(k2, c2))
            tt = DeviceTable(k, k2[:n], c2[:n], bits=self.bits).build_index(direct=(self.path != "symm"))
            if self.path == "symm":
                # (the fingerprint of the freshly unpacked shard is part of the timed call)
                if not fingerprint_verdict(tt.fingerprint(lo, hi, self.seeds), self.group):
                    raise RuntimeError("e2e replica is not symmetric although the resident table was")
                tt.alloc_symm(wlo, whi, self.table.symm_shards)
            else:
                tt.alloc_work(wlo, whi)
            self.scan_on(tt)
            if self.rank == 0:
                h_plot.copy_(tt.plot, non_blocking=True)
            torch.cuda.synchronize()
            return tt

        for _ in range(max(warmup, 1)):
            call()
        dist.barrier(self.group)
        t0 = time.perf_counter()
        for _ in range(steps):
            tt = call()
        dist.barrier(self.group)
        dt = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX, group=self.group)
        same = bool(torch.equal(tt.plot, t.plot)) if steps > 0 else True
        dtv = float(dt.item())
        h2d = torch.tensor([h_rec.numel() + h_idx.numel() * 8], dtype=torch.int64, device=dev)
        dist.all_reduce(h2d, group=self.group)                     # whole job, all ranks
        return {"value": n / dtv, "unit": "k-mers/s", "ms_per_step": dtv * 1e3,
                "h2d_bytes_per_step": int(h2d.item()),
                "d2h_bytes_per_step": int(_lib.PLOT_CELLS * 8),
                "api": "smudgeplot_b200.dist: pinned shard records -> H2D -> hm_k_unpack_records -> shard exchange "
                       "(NCCL) -> hm_k_build_bucket_index -> " +
                       ("fingerprint -> runscan -> all-gather(Bloom) -> resolve" if self.path == "symm" else
                        "filter -> pass1 -> deg exchange -> pass2") + " -> all-reduce(plot) -> D2H",
                "plot_matches_resident_scan": same, "exchange": self.exchange}
