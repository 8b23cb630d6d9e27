"""Host-side micro-benchmark of the stored-PSF reader: the reference's dense files (np.load of 131 200 B + crop,
transforms.py:301-309) against the packed sparse bank (detectinblur_b200/psf_bank.py).  CPU only; PSFs are synthetic
sparse canvases with the bank's tap statistics.

    python tools/bench_bank.py [--n 2000] [--dir /tmp/dib_bank_bench]
"""
import argparse
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detectinblur_b200 import psf_bank  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--dir", default="/tmp/dib_bank_bench")
    args = ap.parse_args()
    shutil.rmtree(args.dir, ignore_errors=True)
    folder = os.path.join(args.dir, "P1E2")
    os.makedirs(folder)
    rng = np.random.default_rng(0)
    for i in range(args.n):
        c = np.zeros((256, 256), np.float16)
        t = int(rng.integers(39, 58))                                    # P1 x E2 tap counts
        ys = np.clip(128 + np.cumsum(rng.integers(-1, 2, t)), 64, 191)
        xs = np.clip(128 + np.cumsum(rng.integers(-1, 2, t)), 64, 191)
        c[ys, xs] = (rng.random(t) / t).astype(np.float16)
        with open(os.path.join(folder, "I%06d" % i), "wb") as f:
            np.save(f, c)
    dense_bytes = sum(os.path.getsize(os.path.join(folder, f)) for f in os.listdir(folder))
    order = rng.integers(0, args.n, 4 * args.n)

    def timed(directory):
        psf_bank._banks.clear()
        psf_bank.load_stored_psf(directory, 1, 2, 0)
        t0 = time.perf_counter()
        acc = 0.0
        for i in order:
            acc += float(psf_bank.load_stored_psf(directory, 1, 2, int(i))[64, 64])
        return len(order) / (time.perf_counter() - t0)

    dense_rate = timed(args.dir)
    packed_dir = args.dir + "_packed"
    shutil.rmtree(packed_dir, ignore_errors=True)
    shutil.copytree(args.dir, packed_dir)
    psf_bank.pack_psf_bank(packed_dir, remove_dense=True)
    pack_bytes = os.path.getsize(os.path.join(packed_dir, "P1E2.dibpack"))
    packed_rate = timed(packed_dir)
    print({"psfs": args.n, "dense_bytes_per_psf": dense_bytes // args.n, "packed_bytes_per_psf": round(pack_bytes / args.n, 1),
           "dense_loads_per_s": round(dense_rate), "packed_loads_per_s": round(packed_rate),
           "note": "page-cache-warm files; one core"})
    shutil.rmtree(args.dir, ignore_errors=True)
    shutil.rmtree(packed_dir, ignore_errors=True)


if __name__ == "__main__":
    main()
