#!/usr/bin/env python
"""Ground truth for the golden cases: tests/golden/truth/<case>.npz.

Runs the extended-precision build of the oracle (oracle/_build/liboracle_quad.so:
the reference algorithm with every scalar in _Float128, FP64 inputs and outputs)
on the effective inputs stored in tests/golden/<case>.npz and stores its outputs.
Needs neither /root/reference nor a GPU.  The parity tests compare both the CUDA
path and the FP64 oracle with these files (tests/parity.py).

Usage: python tests/golden/make_truth.py [case ...]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TRUTH_DIR = os.path.join(HERE, "truth")


def one(case):
    import golden_io
    import oracle_lib
    from spartacus_surface_b200.driver.spartacus_surface_driver import run_radsurf
    t0 = time.time()
    r, _ = golden_io.load_case(case, legendre_gauss_init=oracle_lib.legendre_gauss_init)
    run_radsurf(r, oracle_lib.make_solver(quad=True, nthreads=1))
    out = {}
    for name, fields in golden_io.outputs_of(r).items():
        for k, v in fields.items():
            out[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(TRUTH_DIR, case), **out)
    return case, time.time() - t0


if __name__ == "__main__":
    import golden_io
    os.makedirs(TRUTH_DIR, exist_ok=True)
    cases = [c if c.endswith(".npz") else c + ".npz" for c in sys.argv[1:]] or golden_io.list_cases()
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        for case, dt in pool.imap_unordered(one, cases):
            print(f"{case}: {dt:.1f} s", flush=True)
