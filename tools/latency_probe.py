"""sample() latency at small batches (device-resident loop), for the launch-overhead / CUDA-graph question."""
import os, sys, time, json
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from layoutdm_b200 import Engine, Vocab, timestep_plan
from layoutdm_b200.synthetic import random_state_dict
vocab = Vocab.for_dataset("rico25")
eng = Engine.from_state_dict(random_state_dict(vocab), vocab)
cfg = {"name": "random", "temperature": 1.0}
out = {}
for B, T_eval in ((1, 100), (8, 50), (8, 100), (64, 100), (256, 100)):
    plan = timestep_plan(100, T_eval)
    for _ in range(2): eng.sample_loop(B, plan, cfg, seed=1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 3
    for i in range(n): eng.sample_loop(B, plan, cfg, seed=2 + i)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    out[f"B{B}_T{T_eval}"] = {"ms_per_call": round(dt * 1e3, 2), "us_per_step": round(dt * 1e6 / T_eval, 1), "layouts_per_s": round(B / dt, 1)}
print(json.dumps({"pdl": os.environ.get("LDM_PDL", "1"), "graph": os.environ.get("LDM_GRAPH", "1"), **out}))
