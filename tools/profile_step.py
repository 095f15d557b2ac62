"""Short driver for ncu: a few denoising steps at one of the BASELINE.json workloads.
    LDM_GRAPH=0 ncu ... python tools/profile_step.py [--config 1|2|3] [--batch N] [--steps 3] [--dtype fp16]
  config 1: rico25 unconditional, T=100, batch 1024 (the bench line)      config 2: publaynet cond=c, top_p=0.9, batch 1024
  config 3: rico25 cond=refinement (logit masking), T=200, batch 4096
Launch order per step: embed, 4 x (qkv, attention, outproj_ln, ff1, ff2_ln), head, posterior_sample = 23 kernels."""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from layoutdm_b200 import Engine, Vocab, timestep_plan  # noqa: E402
from layoutdm_b200.synthetic import random_state_dict, synthetic_cond  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=1)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dtype", default="fp16")
a = ap.parse_args()
ds, T, B, cfg, ctype = {1: ("rico25", 100, 1024, {"name": "random", "temperature": 1.0}, None),
                        2: ("publaynet", 100, 1024, {"name": "top_p", "temperature": 1.0, "top_p": 0.9}, "c"),
                        3: ("rico25", 200, 4096, {"name": "random", "temperature": 1.0}, "refinement")}[a.config]
B = a.batch or B
vocab = Vocab.for_dataset(ds)
eng = Engine.from_state_dict(random_state_dict(vocab, num_timesteps=T), vocab, num_timesteps=T, operand_dtype=a.dtype)
cond = None
if ctype:
    cond = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in synthetic_cond(vocab, B, ctype).items()}
plan = timestep_plan(T, T)[: a.steps]
ids = eng.sample_loop(B, plan, cfg, cond=cond, seed=1, ids_init=cond["seq"] if cond else None)
torch.cuda.synchronize()
print("done", int(ids.max()), eng.launch_count)
