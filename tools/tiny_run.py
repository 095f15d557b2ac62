"""Smallest end-to-end run (for compute-sanitizer): B layouts, a few denoising steps, every sampling mode once."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from layoutdm_b200 import Engine, Vocab, timestep_plan
from layoutdm_b200.synthetic import random_state_dict
B = int(os.environ.get("B", "3")); steps = int(os.environ.get("STEPS", "2"))
vocab = Vocab.for_dataset("rico25")
eng = Engine.from_state_dict(random_state_dict(vocab), vocab)
plan = timestep_plan(100, 100)[:steps]
for cfg in ({"name": "random", "temperature": 1.0}, {"name": "top_p", "temperature": 1.0, "top_p": 0.9}, {"name": "deterministic"}):
    ids = eng.sample_loop(B, plan, cfg, seed=1)
    torch.cuda.synchronize()
    assert int(ids.min()) >= 0 and int(ids.max()) < vocab.C
print("tiny run ok", eng.launch_count)
