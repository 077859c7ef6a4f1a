#!/usr/bin/env python3
"""Random-argument stress beyond the seeds the test suite pins: the three random-problem generators of tests/test_gpu_parity.py
(1-d across the round-2 paths, 1-d / 3-d, 2-d) for seeds 48 ... 419 against F.conv* in float64 (tolerance 1e-4)."""
import sys, traceback
sys.path.insert(0, '/root/repo')
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from tests.test_gpu_parity import test_random_1d_problems_across_the_round_2_paths as t1, test_random_1d_and_3d_problems_match_torch as t2, test_random_2d_problems_match_torch as t3
bad = 0
for seed in range(48, 420):
    for fn in (t1, t2, t3):
        try:
            fn(seed)
        except Exception as e:
            bad += 1
            print("FAIL", fn.__name__, seed, repr(e)[:300], flush=True)
    if seed % 50 == 0:
        print("seed", seed, "failures so far", bad, flush=True)
print("done, failures:", bad)
