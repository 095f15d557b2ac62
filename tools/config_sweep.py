"""Device-resident sample() throughput at the five BASELINE.json configs (synthetic weights / conditions), 1 GPU.
Config 4 (8 x 1024 over 8 GPUs) is run as its per-GPU shard.  Prints one JSON line."""
import json, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from layoutdm_b200 import Engine, Vocab, timestep_plan
from layoutdm_b200.synthetic import random_state_dict, synthetic_cond

CONFIGS = [
    ("cfg0 rico25 uncond T_eval=50 B=8 random", "rico25", 100, 50, 8, {"name": "random", "temperature": 1.0}, None),
    ("cfg1 rico25 uncond T=100 B=1024 random", "rico25", 100, 100, 1024, {"name": "random", "temperature": 1.0}, None),
    ("cfg2 publaynet cond=c T=100 B=1024 top_p=0.9", "publaynet", 100, 100, 1024, {"name": "top_p", "temperature": 1.0, "top_p": 0.9}, "c"),
    ("cfg3 rico25 refinement T=200 B=4096 random", "rico25", 200, 200, 4096, {"name": "random", "temperature": 1.0}, "refinement"),
    ("cfg4 rico25 uncond T=100 B=1024 (one of 8 shards)", "rico25", 100, 100, 1024, {"name": "random", "temperature": 1.0}, None),
]
out = {}
for name, ds, T, T_eval, B, cfg, ctype in CONFIGS:
    vocab = Vocab.for_dataset(ds)
    eng = Engine.from_state_dict(random_state_dict(vocab, num_timesteps=T), vocab, num_timesteps=T)
    cond = None
    if ctype:
        cond = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in synthetic_cond(vocab, B, ctype).items()}
    plan = timestep_plan(T, T_eval)
    ids0 = cond["seq"] if cond else None
    eng.sample_loop(B, plan, cfg, cond=cond, seed=1, ids_init=ids0)
    torch.cuda.synchronize()
    n = 2 if B >= 1024 else 5
    t0 = time.perf_counter()
    for i in range(n):
        ids = eng.sample_loop(B, plan, cfg, cond=cond, seed=2 + i, ids_init=ids0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    if cond is not None:                                   # strong conditioning must be reproduced exactly
        assert torch.equal(ids[cond["mask"]], cond["seq"][cond["mask"]])
    assert int(ids.max()) < vocab.C - 1                   # no MASK left after the last step
    out[name] = {"ms_per_call": round(dt * 1e3, 1), "layouts_per_s": round(B / dt, 1), "ms_per_step": round(dt * 1e3 / T_eval, 3)}
    del eng
    torch.cuda.empty_cache()
print(json.dumps(out))
