"""CPU pin of the identity the strand-symmetric scan (csrc/hm_symm.cu) rests on, against the oracle:
on a table that holds rc(x) with count(x) for every x,

    deg(x) = H(x) + U(rc x)      H = partners at positions >= k/2 (all found inside x's run)
                                 U = partners at positions >= k - k/2

and the plot is what oracle_util.partial_runscan / partial_resolve (the two kernels' contracts restated in
Python) produce: isolated high pairs counted twice, or once when they differ at the middle base of an odd k.
No GPU needed; the GPU tests check the kernels against the same oracle."""
import numpy as np
import pytest

import oracle_util as ou
from smudgeplot_b200 import fastk
from tools import synth


def _symmetric_table(k, n0, cmax, seed, dense=False):
    rng = np.random.default_rng(seed)
    space = 4 ** k
    if dense or space < (1 << 40):
        pick = rng.choice(space, size=min(n0, space), replace=False).astype(np.uint64)
    elif k < 32:
        pick = rng.integers(0, space, size=n0, dtype=np.int64).astype(np.uint64)
    else:                                                                # 4^32 does not fit an int64 bound
        pick = (rng.integers(0, 1 << 62, size=n0, dtype=np.int64).astype(np.uint64) << np.uint64(2)) | \
               rng.integers(0, 4, size=n0).astype(np.uint64)
    vals = pick << np.uint64(64 - 2 * k)
    # plant one-substitution partners so that isolated pairs exist at every kind of position
    pos = rng.integers(0, k, size=vals.size)
    mate = vals ^ (rng.integers(1, 4, size=vals.size).astype(np.uint64) << (np.uint64(62) - np.uint64(2) * pos.astype(np.uint64)))
    allv = np.concatenate([vals, mate[: vals.size // 2]])
    rc = np.array([ou._rc(int(x), k) for x in allv.tolist()], dtype=np.uint64)
    keys = np.unique(np.concatenate([allv, rc]))
    canon = np.minimum(keys, np.array([ou._rc(int(x), k) for x in keys.tolist()], dtype=np.uint64))
    _, inv = np.unique(canon, return_inverse=True)
    cnt = rng.integers(1, cmax + 1, size=inv.max() + 1).astype(np.uint16)[inv]
    return keys, cnt


@pytest.mark.parametrize("k,n0,cmax,seed", [(21, 1500, 40, 1), (31, 1500, 40, 2), (12, 1200, 40, 3), (32, 1000, 700, 4),
                                            (5, 300, 40, 5), (4, 100, 520, 6), (7, 2000, 6, 7), (17, 1500, 520, 8)])
def test_degree_identity_and_symmetric_plot_equal_the_oracle(k, n0, cmax, seed):
    keys, cnt = _symmetric_table(k, n0, cmax, seed, dense=(k <= 7))
    n = len(keys)
    want_plot, want_deg = ou.oracle_scan(fastk.keys_u64_to_bytes(keys, k), cnt, k)
    pos_of = {int(x): i for i, x in enumerate(keys.tolist())}
    Pr, pup = k // 2, k - k // 2
    H = np.zeros(n, dtype=np.int64)
    U = np.zeros(n, dtype=np.int64)
    for i, x in enumerate(keys.tolist()):
        for p in range(Pr, k):
            sh = 62 - 2 * p
            b = (x >> sh) & 3
            for alt in range(4):
                j = pos_of.get((x & ~(3 << sh)) | (alt << sh)) if alt != b else None
                if j is not None and int(cnt[i]) + int(cnt[j]) <= ou.SMAX:
                    H[i] += 1
                    U[i] += (p >= pup)
    r = np.array([pos_of[ou._rc(int(x), k)] for x in keys.tolist()])
    assert np.array_equal(cnt[r], cnt)                                   # the table is symmetric
    assert np.array_equal((H + U[r]) & 0xFF, want_deg)                   # the identity
    assert np.array_equal(want_deg[r], want_deg)                         # deg(rc x) = deg(x)
    for seg_bits in (1 << 20, 61):                                       # a roomy filter and one full of false hits
        seg, cand = ou.partial_runscan(keys, cnt, k, 0, n, seg_bits)
        plot = ou.partial_resolve(keys, cnt, k, cand, [seg], [int(keys[0])])
        assert np.array_equal(plot, want_plot)
    assert want_plot.sum() > 0 or k <= 7                                 # (crowded tiny-k tables may have no isolated pair)


def test_a_table_that_passes_the_one_kmer_probe_need_not_satisfy_the_identity():
    """why the scan path is chosen by a whole-table fingerprint and not by examine_table's probe: drop one entry
    (not the probed one) and deg != H + U o rc somewhere, although the reference's probe still says "symmetric" """
    k = 21
    keys, cnt = _symmetric_table(k, 1500, 40, 11)
    want_plot, want_deg = ou.oracle_scan(fastk.keys_u64_to_bytes(keys, k), cnt, k)
    paired = np.nonzero(want_deg == 1)[0]
    victim = int(paired[len(paired) // 2])
    if victim in (1, ):                                                  # the probe looks at entry 1 (PloidyPlot.c:1205)
        victim = int(paired[len(paired) // 2 + 1])
    keep = np.ones(len(keys), dtype=bool)
    keep[victim] = False
    k2, c2 = keys[keep], cnt[keep]
    probe = ou._rc(int(k2[1]), k)
    assert probe in set(k2.tolist())                                     # examine_table's verdict: symmetric
    plot2, _ = ou.oracle_scan(fastk.keys_u64_to_bytes(k2, k), c2, k)     # what the reference computes for it
    seg, cand = ou.partial_runscan(k2, c2, k, 0, len(k2), 1 << 20)
    try:
        got = ou.partial_resolve(k2, c2, k, cand, [seg], [int(k2[0])])
        same = np.array_equal(got, plot2)
    except KeyError:                                                     # rc x is not in the table: the kernel raises
        same = False                                                     #   its "not symmetric" status bit here
    assert not same
