"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in smudgeplot_b200/dist.py: prefix
partition, shard exchange, the incidence all-reduce between the passes and the final plot
all-reduce.  The per-rank compute is stood in for by oracle_util.partial_pass1/2 (test
infrastructure restating the two kernels' contracts); the result must equal the oracle's
single-process plot."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, G, P, HET, COV, L, SEED = 17, 1200, 3, 0.03, 60, 8, 5


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_util as ou
        from smudgeplot_b200 import dist as hd
        from tools import synth
        rng = hd.prefix_partition(world)[rank]
        keys, cnt = synth.synth_table(K, G, P, HET, COV, L, SEED, key_range=rng)
        kf, cf, lo, hi = hd.gather_table(keys, cnt.to(torch.int16))
        ku = kf.numpy().view(np.uint64)
        cn = cf.numpy().view(np.uint16)
        deg_part, up = ou.partial_pass1(ku, cn, K, lo, hi)
        deg = torch.from_numpy(deg_part.copy())
        hd.allreduce_deg(deg)
        plot = torch.from_numpy(ou.partial_pass2(cn, deg.numpy(), up, lo, hi).reshape(-1).copy())
        hd.allreduce_plot(plot)
        q.put((rank, lo, hi, ku.copy(), cn.copy(), deg.numpy().copy(), plot.numpy().copy()))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_host_logic_gloo(world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_util as ou
    from smudgeplot_b200 import fastk
    from tools import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    keys, cnt = synth.synth_table(K, G, P, HET, COV, L, SEED)
    ku = synth.keys_to_u64_numpy(keys)
    cn = cnt.numpy().astype(np.uint16)
    want_plot, want_deg = ou.oracle_scan(fastk.keys_u64_to_bytes(ku, K), cn, K)
    cover = 0
    for rank, lo, hi, k2, c2, deg, plot in res:
        assert np.array_equal(k2, ku) and np.array_equal(c2, cn)       # every rank holds the full sorted table
        assert np.array_equal(deg, want_deg)                           # summed incidence array == Pair
        assert np.array_equal(plot.reshape(want_plot.shape), want_plot)
        assert lo == cover
        cover = hi
    assert cover == len(ku)


def test_prefix_partition_tiles_the_prefix_space():
    from smudgeplot_b200 import dist as hd
    for w in (1, 2, 3, 8):
        parts = hd.prefix_partition(w)
        assert parts[0][0] == 0 and parts[-1][1] == 1 << 24
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def test_balanced_offsets_are_a_monotone_cover_and_favour_heavy_prefixes():
    sys.path.insert(0, ROOT)
    from smudgeplot_b200 import dist as hd
    from tools import synth
    keys, _ = synth.synth_table(21, 60000, 2, 0.01, 40, 4, 3)
    n = keys.numel()
    for w in (1, 2, 3, 4, 8, 16):
        o = hd.balanced_offsets(keys, w)
        assert len(o) == w + 1 and o[0] == 0 and o[-1] == n
        assert all(a <= b for a, b in zip(o, o[1:]))
        sizes = [b - a for a, b in zip(o, o[1:])]
        if w > 1:
            assert max(sizes) < 1.25 * min(sizes)              # a mild correction, not a re-partition
            assert sizes[0] < sizes[-1]                        # 'a...' entries cost more than 't...' ones
    # degenerate inputs
    assert hd.balanced_offsets(keys[:1], 4)[-1] == 1
    assert hd.balanced_offsets(keys[:0], 2) == [0, 0, 0]


# ---- strand-symmetric scan, sharded: run-aligned cuts, fingerprint verdict, segment all-gather ----

def _symm_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_util as ou
        from smudgeplot_b200 import dist as hd
        from tools import synth
        rng = hd.prefix_partition(world)[rank]
        keys, cnt = synth.synth_table(K, G, P, HET, COV, L, SEED, key_range=rng)
        kf, cf, lo, hi = hd.gather_table(keys, cnt.to(torch.int16))
        ku = kf.numpy().view(np.uint64)
        cn = cf.numpy().view(np.uint16)
        # fingerprint verdict from per-rank partial sums (stand-in sums: any additive multiset hash)
        def fp(a, b):
            xs = ku[a:b].astype(object)
            h1 = sum((int(x) * 0x9E3779B97F4A7C15 + int(c)) & (2**64 - 1) for x, c in zip(xs, cn[a:b])) & (2**64 - 1)
            h2 = sum((ou._rc(int(x), K) * 0x9E3779B97F4A7C15 + int(c)) & (2**64 - 1) for x, c in zip(xs, cn[a:b])) & (2**64 - 1)
            s = lambda v: v - 2**64 if v >= 2**63 else v
            return torch.tensor([s(h1), s(h1 ^ 5), s(h2), s(h2 ^ 5)], dtype=torch.int64)
        symmetric = hd.fingerprint_verdict(fp(lo, hi))
        cuts = hd.run_aligned_cuts(kf, K, world)
        a, b = cuts[rank], cuts[rank + 1]
        seg_bits = 257
        seg, cand = ou.partial_runscan(ku, cn, K, a, b, seg_bits)
        segs = torch.zeros((world, seg_bits), dtype=torch.uint8)
        segs[rank] = torch.from_numpy(seg)
        hd.exchange_segments(segs, rank)
        first = [int(ku[min(c, len(ku) - 1)]) for c in cuts[:-1]]
        plot = torch.from_numpy(ou.partial_resolve(ku, cn, K, cand, segs.numpy(), first).reshape(-1).copy())
        hd.allreduce_plot(plot)
        q.put((rank, a, b, symmetric, len(cand), plot.numpy().copy()))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_symmetric_scan_host_logic_gloo(world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_util as ou
    from smudgeplot_b200 import fastk
    from tools import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_symm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    keys, cnt = synth.synth_table(K, G, P, HET, COV, L, SEED)
    ku = synth.keys_to_u64_numpy(keys)
    cn = cnt.numpy().astype(np.uint16)
    want_plot, _ = ou.oracle_scan(fastk.keys_u64_to_bytes(ku, K), cn, K)
    assert want_plot.sum() > 0
    cover = 0
    for rank, a, b, symmetric, ncand, plot in res:
        assert symmetric                                               # job-wide verdict on every rank
        assert np.array_equal(plot.reshape(want_plot.shape), want_plot)
        assert a == cover
        cover = b
    assert cover == len(ku)
    assert sum(r[4] for r in res) > 0


def test_run_aligned_cuts_never_split_a_run():
    sys.path.insert(0, ROOT)
    from smudgeplot_b200 import dist as hd
    from tools import synth
    for k in (13, 21, 31):
        keys, _ = synth.synth_table(k, 30000, 2, 0.02, 40, 4, 3)
        n = keys.numel()
        sh = 64 - 2 * (k // 2)
        pre = (keys >> sh) & ((1 << (64 - sh)) - 1)
        for w in (1, 2, 3, 8, 16):
            c = hd.run_aligned_cuts(keys, k, w)
            assert len(c) == w + 1 and c[0] == 0 and c[-1] == n and all(x <= y for x, y in zip(c, c[1:]))
            for x in c[1:-1]:
                assert x == n or int(pre[x]) != int(pre[x - 1])
    one = torch.zeros(5, dtype=torch.int64)                            # a single run: all cuts collapse to n
    assert hd.run_aligned_cuts(one, 21, 3) == [0, 5, 5, 5]
