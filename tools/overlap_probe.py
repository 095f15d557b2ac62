"""Experiment: two half-batch engines on two streams, each with persistent grids sized for a share of the SMs, against one
full-batch engine -- do the tensor-bound kernels of one half overlap the HBM-bound kernels of the other?
    python tools/overlap_probe.py [--share 76,72]"""
import argparse, json, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from layoutdm_b200 import Engine, Vocab, timestep_plan
from layoutdm_b200.synthetic import random_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--share", default="76,72")
ap.add_argument("--batch", type=int, default=1024)
a = ap.parse_args()
vocab = Vocab.for_dataset("rico25")
sd = random_state_dict(vocab)
plan = timestep_plan(100, 100)
cfg = {"name": "random", "temperature": 1.0}
B = a.batch
out = {}


def timed(fn, n=3):
    fn(); fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


full = Engine.from_state_dict(sd, vocab)
dt = timed(lambda: full.sample_loop(B, plan, cfg, seed=1))
out["one_engine_all_sms"] = {"ms": round(dt * 1e3, 1), "layouts_per_s": round(B / dt, 1)}
full.close()
shares = [int(x) for x in a.share.split(",")]
tot = sum(shares)
sizes = [int(round(B * s / tot / 2)) * 2 for s in shares]
sizes[-1] = B - sum(sizes[:-1])
engs, streams = [], []
for s in shares:
    os.environ["LDM_NUM_SMS"] = str(s)
    engs.append(Engine.from_state_dict(sd, vocab))
    streams.append(torch.cuda.Stream())
os.environ.pop("LDM_NUM_SMS")


def both():
    for e, st, n in zip(engs, streams, sizes):
        with torch.cuda.stream(st):
            e.sample_loop(n, plan, cfg, seed=2)


dt2 = timed(both)
out["split"] = {"shares": shares, "batches": sizes, "ms": round(dt2 * 1e3, 1), "layouts_per_s": round(B / dt2, 1)}
# each part alone on its SM share (no overlap partner): how much does the concurrency cost / gain
for e, st, n, s in zip(engs, streams, sizes, shares):
    def one(e=e, st=st, n=n):
        with torch.cuda.stream(st):
            e.sample_loop(n, plan, cfg, seed=3)
    d = timed(one)
    out[f"alone_{s}sms_B{n}"] = {"ms": round(d * 1e3, 1), "layouts_per_s": round(n / d, 1)}
print(json.dumps(out))
