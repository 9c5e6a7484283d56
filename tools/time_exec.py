#!/usr/bin/env python3
"""Wall clock of the drop-in executable on a seeded table in /dev/shm (T_total of SURVEY.md §8d)."""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402
from smudgeplot_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nels", type=float, default=2e8)
    ap.add_argument("--threads", type=int, nargs="+", default=[4, 16, 64])
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    G = synth.calibrate_G(31, int(a.nels), 2, 0.01, 40, 12)
    keys, cnt = synth.synth_table(31, G, 2, 0.01, 40, 12, 2, device=dev)
    d = tempfile.mkdtemp(prefix="hetmers_exec_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    name = os.path.join(d, "tab")
    synth.write_table(name, 31, keys, cnt, ibyte=3, nparts=4)
    del keys, cnt
    if dev == "cuda":
        torch.cuda.empty_cache()
    env = dict(os.environ, HETMERS_STATS="1", HETMERS_GPUS=str(a.gpus))
    for T in a.threads:
        for rep in range(2):
            out = os.path.join(d, "o")
            if os.path.exists(out + ".smu"):
                os.remove(out + ".smu")
            t0 = time.perf_counter()
            r = subprocess.run([_lib.BIN_PATH, "-e12", f"-T{T}", f"-o{out}", name], input="n\n",
                               capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            print(f"-T{T} rep{rep}: wall {dt:.3f} s rc={r.returncode} {r.stderr.strip()[-400:]}", flush=True)
    import shutil
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
