"""CPU tests: the oracle (oracle/hetmers_oracle.c) against the golden vectors produced by the
unmodified reference binary, against the brute-force definition, and the host-side FastK code."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases
import oracle_util as ou
from smudgeplot_b200 import fastk
from tools import synth


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_reference_smu(name, golden_meta, tmp_path):
    c = golden_meta[name]
    out = str(tmp_path / "o.smu")
    rc, trim, symm, nels = ou.oracle_file(os.path.join(GOLDEN, name, name), c["e"], out)
    assert (rc, trim, symm) == (0, 1, 1)
    assert nels == c["nels"]
    want = open(os.path.join(GOLDEN, name, name + ".smu")).read()
    assert open(out).read() == want
    assert len(want.splitlines()) == c["smu_rows"]


def _golden_pairs(name):
    d = os.path.join(GOLDEN, name)
    pre = name + ".pairs."
    return {f[len(pre):-4]: open(os.path.join(d, f)).read().splitlines()
            for f in sorted(os.listdir(d)) if f.startswith(pre)}


@pytest.mark.parametrize("name", [n for n in golden_cases() if os.path.exists(os.path.join(GOLDEN, n, n + ".sma"))])
def test_oracle_extract_reproduces_reference_pair_lists(name, golden_meta, tmp_path):
    """extract_kmer_pairs (PloidyList.c): the oracle's pair lists == the reference binary's (sorted)"""
    c = golden_meta[name]
    want = _golden_pairs(name)
    assert want and {k: len(v) for k, v in want.items()} == c["pairs"]
    out = str(tmp_path / "ex")
    assert ou.oracle_extract(os.path.join(GOLDEN, name, name), c["e"], os.path.join(GOLDEN, name, name + ".sma"), out) == 0
    assert ou.sorted_pair_files(out) == want
    k = c["k"]
    for lines in want.values():                        # print_het format: k bases + "(x/y)" at one position
        assert all(len(ln) == k + 4 and ln.count("(") == 1 and ln[ln.index("(") + 2] == "/" for ln in lines)


def test_oracle_extract_error_codes(tmp_path):
    sma = tmp_path / "s.sma"
    sma.write_text("covB\tcovA\tfreq\tsmudge\n5\t9\t3\t1A1B\n")
    assert ou.oracle_extract(str(tmp_path / "absent"), 4, str(sma), str(tmp_path / "o")) == 1
    assert ou.oracle_extract(os.path.join(GOLDEN, "dip_k21", "dip_k21"), 4, str(tmp_path / "no.sma"), str(tmp_path / "o")) == 1
    sma.write_text("covB\tcovA\tfreq\tsmudge\n9\t5\t3\t1A1B\n")          # covB > covA: invalid pixel
    assert ou.oracle_extract(os.path.join(GOLDEN, "dip_k21", "dip_k21"), 4, str(sma), str(tmp_path / "o")) == 3


@pytest.mark.parametrize("name,trim,symm", [("untrimmed", 0, 1), ("asymmetric", 1, 0)])
def test_oracle_examine_matches_reference_verdict(name, trim, symm, golden_meta, tmp_path):
    c = golden_meta["_conditioning"][name]
    rc, t, s, _ = ou.oracle_file(os.path.join(GOLDEN, "conditioning", name), c["e"], str(tmp_path / "x.smu"))
    assert rc == 2 and (t, s) == (trim, symm)
    words = {(1, 1): "trimmed and symmetric", (1, 0): "trimmed but not symmetric",
             (0, 1): "untrimmed yet symmetric", (0, 0): "untrimmed and not symmetric"}[(t, s)]
    assert c["verbose"] == ["The input table is " + words]       # what the reference printed
    assert c["returncode"] == 1


def test_oracle_missing_table(tmp_path):
    rc, *_ = ou.oracle_file(str(tmp_path / "nope"), 4, str(tmp_path / "x.smu"))
    assert rc == 1


@pytest.mark.parametrize("k,G,ploidy,het,seed", [(9, 3000, 2, 0.05, 3), (21, 1500, 3, 0.03, 5),
                                                 (32, 1200, 2, 0.02, 8), (5, 400, 2, 0.1, 2)])
def test_oracle_equals_brute_force(k, G, ploidy, het, seed):
    keys, cnt = synth.synth_table(k, G, ploidy, het, 40, 4, seed)
    ku = synth.keys_to_u64_numpy(keys)
    cn = cnt.numpy().astype(np.uint16)
    plot_o, deg_o = ou.oracle_scan(fastk.keys_u64_to_bytes(ku, k), cn, k)
    plot_b, deg_b = ou.brute_force(ku, cn, k)
    assert np.array_equal(deg_o, deg_b)
    assert np.array_equal(plot_o, plot_b)


def test_oracle_smax_fmax_gates():
    # two isolated one-away pairs: (aaaaa,aaaac) with the counts under test and (ggggg,ggggt) = (7,9).
    # 500+500 lands in bin m=500 which the writer drops (i < FMAX); 501+500 exceeds SMAX (no pair).
    k = 5
    g5 = 0b1010101010 << 54
    keys = np.array([0, 1 << 54, g5, g5 + (1 << 54)], dtype=np.uint64)
    for cnts, rows in (([500, 500, 7, 9], "7\t9\t1\n"), ([501, 500, 7, 9], "7\t9\t1\n"),
                       ([499, 500, 7, 9], "7\t9\t1\n499\t500\t1\n")):
        plot, deg = ou.oracle_scan(fastk.keys_u64_to_bytes(keys, k), np.array(cnts, dtype=np.uint16), k)
        pb, db = ou.brute_force(keys, np.array(cnts), k)
        assert np.array_equal(plot, pb) and np.array_equal(deg, db)
        assert ou.smu_text(plot) == rows
    assert plot[999, 499] == 1


def test_empty_and_single_entry_tables():
    for n in (0, 1):
        keys = np.zeros((n, 6), dtype=np.uint8)
        plot, deg = ou.oracle_scan(keys, np.ones(n, dtype=np.uint16), 21)
        assert plot.sum() == 0 and deg.size == n


# ----------------------------------------------------------------------- host FastK code ----

@pytest.mark.parametrize("k,ibyte,nparts", [(21, 1, 1), (31, 2, 3), (32, 3, 2), (11, 1, 4), (12, 3, 1)])
def test_fastk_write_read_roundtrip(k, ibyte, nparts, tmp_path):
    keys, cnt = synth.synth_table(k, 1500, 2, 0.02, 40, 4, 13)
    ku = synth.keys_to_u64_numpy(keys)
    cn = cnt.numpy().astype(np.uint16)
    name = str(tmp_path / "tab")
    fastk.write_ktab(name, k, ku, cn, ibyte=ibyte, nparts=nparts)
    kt = fastk.read_ktab(name + ".ktab")          # suffix optional / accepted
    assert (kt.kmer, kt.ibyte, kt.nparts, kt.nels) == (k, ibyte, nparts, len(ku))
    kb, c2 = fastk.unpack_host(kt)
    assert np.array_equal(fastk.keys_bytes_to_u64(kb), ku)
    assert np.array_equal(c2, cn)
    assert kt.index[-1] == len(ku)
    fastk.remove_ktab(name)
    assert not os.path.exists(name + ".ktab")


def test_fastk_reader_matches_golden_and_oracle_reader(golden_meta):
    # the python reader and the oracle's C reader must agree on every golden table
    for name in golden_cases():
        kt = fastk.read_ktab(os.path.join(GOLDEN, name, name))
        kb, cn = fastk.unpack_host(kt)
        assert kt.nels == golden_meta[name]["nels"]
        kv = np.ascontiguousarray(kb).view(f"S{kb.shape[1]}").reshape(-1)      # byte-string order == table order
        assert np.all(kv[1:] > kv[:-1])
        plot, _ = ou.oracle_scan(kb, cn, kt.kmer)
        assert ou.smu_text(plot) == open(os.path.join(GOLDEN, name, name + ".smu")).read()


def test_fastk_missing_part_and_stub(tmp_path):
    with pytest.raises(FileNotFoundError):
        fastk.read_ktab(str(tmp_path / "absent"))
    keys, cnt = synth.synth_table(21, 600, 2, 0.02, 40, 4, 1)
    name = str(tmp_path / "t")
    fastk.write_ktab(name, 21, synth.keys_to_u64_numpy(keys), cnt.numpy(), ibyte=1, nparts=2)
    os.remove(fastk.part_path(name, 2))
    with pytest.raises(FileNotFoundError):
        fastk.read_ktab(name)


def test_synth_table_is_trimmed_symmetric_and_deterministic():
    a = synth.synth_table(21, 3000, 3, 0.02, 60, 12, 7)
    b = synth.synth_table(21, 3000, 3, 0.02, 60, 12, 7)
    assert all((x == y).all() for x, y in zip(a, b))
    keys, cnt = a
    assert int(cnt.min()) >= 12 and int(cnt.max()) <= 32767
    ku = synth.keys_to_u64_numpy(keys)
    assert (ku[1:] > ku[:-1]).all()
    rc = synth.keys_to_u64_numpy(synth.revcomp_left(keys, 21))
    pos = np.searchsorted(ku, rc)
    assert (ku[pos] == rc).all()                    # symmetric
    assert (cnt.numpy()[pos] == cnt.numpy()).all()  # with equal counts
    # shards by 24-bit prefix tile the table
    lo = synth.synth_table(21, 3000, 3, 0.02, 60, 12, 7, key_range=(0, 1 << 23))
    hi = synth.synth_table(21, 3000, 3, 0.02, 60, 12, 7, key_range=(1 << 23, 1 << 24))
    assert np.array_equal(np.concatenate([synth.keys_to_u64_numpy(lo[0]), synth.keys_to_u64_numpy(hi[0])]), ku)


# ------------------------------------------------------------------ randomised small tables ----

def _random_table(rng, k, n, cmax):
    space = 4 ** k
    n = min(n, space)
    vals = np.sort(rng.choice(space, size=n, replace=False).astype(np.uint64))
    keys = vals << np.uint64(64 - 2 * k)
    cnt = rng.integers(1, cmax + 1, size=n).astype(np.uint16)
    return keys, cnt


@pytest.mark.parametrize("seed", range(12))
def test_oracle_equals_brute_force_on_random_dense_tables(seed):
    """arbitrary (not genome-like, not symmetric) tables: dense neighbourhoods, counts around the
    SMAX = 1000 / FMAX = 500 gates, tiny k -- the oracle's trie/merge restatement against the
    ten-line definition of SURVEY.md Appendix B"""
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.integers(1, 9))
    n = int(rng.integers(2, 600))
    cmax = int(rng.choice([6, 40, 520, 700]))
    keys, cnt = _random_table(rng, k, n, cmax)
    plot_o, deg_o = ou.oracle_scan(fastk.keys_u64_to_bytes(keys, k), cnt, k)
    plot_b, deg_b = ou.brute_force(keys, cnt, k)
    assert np.array_equal(deg_o, deg_b)
    assert np.array_equal(plot_o, plot_b)
    # and the range-restricted kernel contracts (multi-GPU host-logic tests) compose to the same plot
    cuts = sorted({0, len(keys)} | set(int(v) for v in rng.integers(0, len(keys) + 1, size=2)))
    deg = np.zeros(len(keys), dtype=np.uint8)
    ups = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        d, up = ou.partial_pass1(keys, cnt, k, lo, hi)
        deg += d
        ups.append((lo, hi, up))
    plot = sum(ou.partial_pass2(cnt, deg, up, lo, hi) for lo, hi, up in ups)
    assert np.array_equal(deg, deg_b) and np.array_equal(plot, plot_b)
