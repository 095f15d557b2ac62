"""Short driver for ncu: a few denoising steps at the bench workload (rico25 uncond, B=1024 by default).
    ncu ... python tools/profile_step.py [--batch 1024] [--steps 3] [--dtype fp16]
Launch order per step: embed, 4 x (qkv, attention, outproj_ln, ff1, ff2_ln), head, posterior_sample = 23 kernels."""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from layoutdm_b200 import Engine, Vocab, timestep_plan  # noqa: E402
from layoutdm_b200.synthetic import random_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dtype", default="fp16")
a = ap.parse_args()
vocab = Vocab.for_dataset("rico25")
eng = Engine.from_state_dict(random_state_dict(vocab), vocab, operand_dtype=a.dtype)
plan = timestep_plan(100, 100)[: a.steps]
ids = eng.sample_loop(a.batch, plan, {"name": "random", "temperature": 1.0}, seed=1)
torch.cuda.synchronize()
print("done", int(ids.max()), eng.launch_count)
