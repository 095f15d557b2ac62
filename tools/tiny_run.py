"""Smallest end-to-end run (for compute-sanitizer): B layouts, a few denoising steps, every sampling mode once, the graph-replayed
loop twice, cond = c / refinement / relation, the corruption + training-side entry points, decode and cond construction."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from layoutdm_b200 import Engine, Vocab, timestep_plan
from layoutdm_b200.synthetic import random_state_dict, synthetic_cond
B = int(os.environ.get("B", "3")); steps = int(os.environ.get("STEPS", "2"))
vocab = Vocab.for_dataset("rico25")
eng = Engine.from_state_dict(random_state_dict(vocab), vocab)
plan = timestep_plan(100, 100)[:steps]
for cfg in ({"name": "random", "temperature": 1.0}, {"name": "top_p", "temperature": 1.0, "top_p": 0.9}, {"name": "deterministic"},
            {"name": "gumbel", "temperature": 1.0}, {"name": "top_k", "temperature": 1.0, "top_k": 3}):
    for rep in range(2):                                   # second call replays the captured graph
        ids = eng.sample_loop(B, plan, cfg, seed=1)
    torch.cuda.synchronize()
    assert int(ids.min()) >= 0 and int(ids.max()) < vocab.C
rnd = {"name": "random", "temperature": 1.0}
for ctype in ("c", "refinement"):
    cond = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in synthetic_cond(vocab, B, ctype).items()}
    ids = eng.sample_loop(B, plan, rnd, cond=cond, seed=2, ids_init=cond["seq"])
    assert torch.equal(ids[cond["mask"]], cond["seq"][cond["mask"]])
cond = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in synthetic_cond(vocab, B, "c").items()}
adj = torch.zeros(B, 26, 26, dtype=torch.int32)
adj[:, 0, 1] = (1 << 1) | (1 << 6); adj[:, 1, 2] = (1 << 2) | (1 << 5); adj[:, 2, 3] = (1 << 3) | (1 << 9)
cond.update(type="relation", rel_adj=adj.cuda(), rel_lambda=3e4, rel_num_update=3)
ids, trace = eng.sample_loop(B, timestep_plan(100, 4), rnd, cond=cond, seed=3, ids_init=cond["seq"], trace=True)
x0 = ids.clone()
t = torch.tensor([0, 50, 99][:B] + [7] * max(0, B - 3)).cuda()
xt = eng.q_sample(x0, t, seed=4)
lx0 = eng.predict_start(xt, t)
post = eng.q_posterior(lx0, xt, t)
qp = eng.q_pred(lx0, t); q1 = eng.q_pred_one_timestep(lx0, t); ga = eng.gumbel_argmax(qp, seed=5)
r = eng.vb_terms(x0, xt, t, want_log_model_prob=True, want_recon_ids=True)
lay = eng.decode(ids)
c2 = eng.cond_from_layouts(lay["label"], lay["bbox"], lay["mask"], "cwh")
torch.cuda.synchronize()
assert torch.isfinite(r["kl"]).all() and torch.isfinite(post).all() and c2["seq"].shape == ids.shape
print("tiny run ok", eng.launch_count)
