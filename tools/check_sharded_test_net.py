#!/usr/bin/env python
"""GPU-box check of the dataset loop (the reference's tools/test_net.py flow) in its three launch modes on one synthetic imdb:

    single process, batch 1   |   torchrun, 2 ranks over NCCL (image i -> rank i mod 2, one all-gather per step)   |   single, --batch 2

and comparison of the three detections.pkl files: sharded == single bit for bit (same plans, same kernels), batched within the
summation-order tolerance (split-K layers see another M).  Prints a report; exit code 0 iff everything agrees.
    python tools/check_sharded_test_net.py [--imdb synthetic_6_21] [--net res50]"""
import argparse
import os
import pickle
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd, env=None):
    print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-3000:], r.stderr[-3000:])
        raise SystemExit("command failed")
    return r.stdout


def load(tag, imdb):
    path = os.path.join(ROOT, "output", "default", imdb, "default", tag, "detections.pkl")
    with open(path, "rb") as f:
        return pickle.load(f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--imdb", default="synthetic_6_21")
    ap.add_argument("--net", default="res50")
    a = ap.parse_args()
    tool = os.path.join("tools", "test_net.py")
    base = [tool, "--imdb", a.imdb, "--net", a.net]
    run([sys.executable] + base + ["--tag", "single"])
    run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29533"] + base + ["--tag", "sharded"])
    run([sys.executable] + base + ["--tag", "batch2", "--batch", "2"])
    single, sharded, batched = load("single", a.imdb), load("sharded", a.imdb), load("batch2", a.imdb)
    ncls, nimg = len(single), len(single[0])
    total = sum(len(single[j][i]) for j in range(1, ncls) for i in range(nimg))
    same = all(np.array_equal(np.asarray(single[j][i], np.float32).reshape(-1, 5), np.asarray(sharded[j][i], np.float32).reshape(-1, 5))
               for j in range(1, ncls) for i in range(nimg))
    worst, mismatched = 0.0, 0
    for j in range(1, ncls):
        for i in range(nimg):
            x = np.asarray(single[j][i], np.float32).reshape(-1, 5); y = np.asarray(batched[j][i], np.float32).reshape(-1, 5)
            if x.shape != y.shape:
                mismatched += abs(x.shape[0] - y.shape[0])
                continue
            if x.size:
                worst = max(worst, float(np.abs(x - y).max()))
    print("imdb %s, net %s: %d images, %d detections in the single-process run" % (a.imdb, a.net, nimg, total))
    print("2-rank NCCL run == single-process run, bit for bit: %s" % same)
    print("--batch 2 run vs single: %d detections differ in count, max |difference| on the rest %.3g" % (mismatched, worst))
    ok = same and mismatched <= max(2, total // 50) and worst < 5e-2
    print("OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
