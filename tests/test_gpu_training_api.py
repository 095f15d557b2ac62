"""GPU (-m gpu): the training-side API of SURVEY 8b / 8f-3 through the C ABI -- predict_start / q_posterior / q_pred on log
tensors with PER-LAYOUT timesteps and the loss terms of `forward` (ldm_vb_terms) -- against the oracle restatement that
tests/test_oracle_vs_reference.py pins to the unmodified reference (both q_types).  Tolerance 1e-4 on log-probabilities / loss
terms (fp32 elementwise math), 1e-3 on logits (16-bit tensor-core operands) at the reference's weight scale."""
import pytest
import torch

from oracle import layoutdm_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4
T_CASES = [0, 1, 50, 99, 37, 0, 98, 10, 9]


def setup(q_type, scale=1.0):
    from layoutdm_b200 import Engine, Vocab
    vo, spec = O.RICO25, O.ModelSpec()
    sd = O.make_weights(vo, spec, seed=5, scale=scale)
    eng = Engine.from_state_dict(sd, Vocab.for_dataset("rico25"), num_timesteps=spec.T, q_type=q_type)
    return eng, sd, vo, spec, O.group_schedules(spec.T, vo, q_type)


def inputs(vo, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.empty(B, vo.S, dtype=torch.long)
    for a in range(5):
        ids = torch.tensor(vo.group_full_ids(a)[:-1])
        x0[:, a::5] = ids[torch.randint(0, len(ids), (B, 25), generator=g)]
    t = torch.tensor((T_CASES * ((B + len(T_CASES) - 1) // len(T_CASES)))[:B])
    xt = O.q_sample_ids(x0, t, 100, vo, O.group_schedules(100, vo), O.uniforms(3, 0, 2, 0, B, vo.S, vo.C))
    return x0, xt, t, g


@pytest.mark.parametrize("q_type", ["constrained", "vanilla"])
def test_q_posterior_and_q_pred_match_oracle(q_type):
    eng, sd, vo, spec, scheds = setup(q_type)
    B = 18
    x0, xt, t, g = inputs(vo, B)
    lx = torch.log_softmax(torch.randn(B, vo.S, vo.C, generator=g) * 2.0, dim=-1).clamp(-70.0, 0.0)
    got = eng.q_posterior(lx.cuda(), xt.cuda(), t.cuda()).cpu()
    want = O.q_posterior(lx, xt, t, spec.T, vo, scheds, q_type)
    assert (got - want).abs().max() < TOL
    one = O.index_to_log_onehot(x0, vo.C)                                     # log_true_prob path: one-hot x0
    got1 = eng.q_posterior(one.cuda(), xt.cuda(), t.cuda()).cpu()
    assert (got1 - O.q_posterior(one, xt, t, spec.T, vo, scheds, q_type)).abs().max() < TOL
    tq = t.clone(); tq[0] = -1
    gp = eng.q_pred(lx.cuda(), tq.cuda()).cpu()
    assert (gp - O.q_pred_full(lx, tq, spec.T, vo, scheds, q_type)).abs().max() < TOL
    g1 = eng.q_pred_one_timestep(lx.cuda(), t.cuda()).cpu()
    assert (g1 - O.q_pred_one_timestep_full(lx, t, spec.T, vo, scheds, q_type)).abs().max() < TOL
    # log_sample_categorical (gumbel argmax): bit-exact vs the oracle under the shared noise, and == ldm_q_sample when fed q_pred(one-hot x0)
    ids = eng.gumbel_argmax(lx.cuda(), seed=7).cpu()
    assert torch.equal(ids, O.gumbel_argmax(lx, O.uniforms(7, 0, 2, 0, B, vo.S, vo.C)))
    if q_type == "constrained":
        qp = eng.q_pred(one.cuda(), t.cuda())
        qp = torch.where(qp > O.LOG_EPS + 1e-3, qp, torch.full_like(qp, float("-inf")))      # classes outside the token's group are impossible
        assert torch.equal(eng.gumbel_argmax(qp, seed=21), eng.q_sample(x0.cuda(), t.cuda(), seed=21))


@pytest.mark.parametrize("q_type,B", [("constrained", 18), ("constrained", 301), ("vanilla", 9)])
def test_predict_start_and_vb_terms_per_layout_timesteps(q_type, B):
    """denoiser with per-layout AdaLN rows (embed + the FF2 epilogue reload their (scale, shift) per layout) + the loss terms"""
    eng, sd, vo, spec, scheds = setup(q_type)
    x0, xt, t, g = inputs(vo, B, seed=B)
    lx0, logits = eng.predict_start(xt.cuda(), t.cuda(), want_logits=True)
    lx0, logits = lx0.cpu(), logits.cpu()
    with torch.no_grad():
        ref = torch.cat([O.denoiser_forward(sd, xt[i:i + 128], t[i:i + 128], vo, spec) for i in range(0, B, 128)])
    err = (logits - ref).abs().amax(dim=(1, 2))
    print(f"{q_type} B={B}: logits max-abs error {err.max():.2e} (per-layout timesteps {sorted(set(t.tolist()))})")
    assert err.max() < 1e-3, f"layouts off: {(err >= 1e-3).nonzero().flatten().tolist()[:8]} t={t[(err >= 1e-3)].tolist()[:8]}"
    assert (lx0 - O.predict_start(logits)).abs().max() < TOL
    # same timestep for everyone == the sampling path's scalar-t denoiser, bit for bit
    t_same = torch.full((B,), 42)
    _, lg_a = eng.predict_start(xt.cuda(), t_same.cuda(), want_logits=True)
    _, lg_b, _ = eng.step(xt.cuda(), 42, 42, {"name": "deterministic"}, want_logits=True)
    assert torch.equal(lg_a, lg_b)
    r = eng.vb_terms(x0.cuda(), xt.cuda(), t.cuda(), (1.0, 1.0), want_log_model_prob=True, want_recon_ids=True)
    w = O.vb_terms(logits, x0, xt, t, spec.T, vo, scheds, q_type)
    assert (r["log_model_prob"].cpu() - w["log_model_prob"]).abs().max() < TOL
    for k in ("kl", "decoder_nll", "kl_aux"):
        d = (r[k].cpu() - w[k]).abs()
        assert (d <= TOL * (1.0 + w[k].abs())).all(), f"{k}: {d.max():.3e}"
    assert torch.equal(r["x0_recon"].cpu(), w["log_x0_recon"].argmax(-1))
    assert (r["xt_1_recon"].cpu() != w["log_model_prob"].argmax(-1)).float().mean() < 1e-3     # ties between clamped entries
    r2 = eng.vb_terms(x0.cuda(), xt.cuda(), t.cuda(), (2.0, 0.5), want_aux=False)
    w2 = O.vb_terms(logits, x0, xt, t, spec.T, vo, scheds, q_type, mask_weight=(2.0, 0.5))
    assert r2["kl_aux"] is None and (r2["kl"].cpu() - w2["kl"]).abs().max() < TOL * (1.0 + w2["kl"].abs().max())


def test_reference_class_api_training_side():
    """FusedMaskAndReplaceDiffusion keeps the reference's signatures: q_pred(log_x_start, t, key) / q_posterior / predict_start on
    (B,C,S) tensors, q_sample on partial one-hots, forward(x) -> (outputs, losses)"""
    from layoutdm_b200 import LayoutDMB200
    vo, spec = O.RICO25, O.ModelSpec()
    sd = O.make_weights(vo, spec, seed=5, scale=1.0)
    core = LayoutDMB200.from_state_dict(sd, dataset="rico25", num_timesteps=100).model
    scheds = O.group_schedules(100, vo)
    B = 6
    x0, xt, t, g = inputs(vo, B, seed=1)
    log_xt = O.index_to_log_onehot(xt, vo.C).permute(0, 2, 1).cuda()
    lx0 = core.predict_start(log_xt, t.cuda())
    assert lx0.shape == (B, vo.C, vo.S)
    post = core.q_posterior(lx0, log_xt, t.cuda())
    want = O.q_posterior(lx0.permute(0, 2, 1).cpu(), xt, t, 100, vo, scheds)
    assert (post.permute(0, 2, 1).cpu() - want).abs().max() < TOL
    for a, key in enumerate("cxywh"):
        idx = torch.tensor(vo.group_full_ids(a))
        part = torch.log_softmax(torch.randn(B, len(idx), 25, generator=g), dim=1)
        got = core.q_pred(part.cuda(), t.cuda(), key).cpu()
        full = torch.full((B, vo.S, vo.C), O.LOG_EPS)
        full[:, a::5, idx] = part.permute(0, 2, 1)
        want_p = O.q_pred_full(full, t, 100, vo, scheds)[:, a::5][..., idx].permute(0, 2, 1)
        assert got.shape == part.shape and (got - want_p).abs().max() < TOL
        one = torch.log(torch.nn.functional.one_hot(torch.randint(0, len(idx) - 2, (B, 25), generator=g), len(idx)).permute(0, 2, 1).float().clamp(min=1e-30))
        xs = core.q_sample(one.cuda(), t.cuda(), key, seed=3)
        assert xs.shape == one.shape and torch.allclose(xs.exp().sum(1), torch.ones(B, 25, device=xs.device), atol=1e-6)
        one_t = core.q_pred_one_timestep(part.cuda(), t.cuda(), key).cpu()
        want_1 = O.q_pred_one_timestep_full(full, t, 100, vo, scheds)[:, a::5][..., idx].permute(0, 2, 1)
        assert (one_t - want_1).abs().max() < TOL
        ls = core.log_sample_categorical(part.cuda(), key, seed=5)
        assert ls.shape == part.shape and torch.allclose(ls.exp().sum(1), torch.ones(B, 25, device=ls.device), atol=1e-6)
    lg = torch.randn(B, vo.C, vo.S, generator=g)
    drawn = core.sample_logits(lg.cuda(), {"name": "deterministic"})
    assert drawn.shape == (B, 1, vo.S) and torch.equal(drawn[:, 0].cpu(), lg.argmax(1))
    pt = torch.full((B,), 0.01)
    outputs, losses = core.forward(x0.cuda(), is_train=True, t=t, pt=pt, seed=11)
    assert outputs["probs"].shape == (B, vo.C, vo.S) and torch.isfinite(losses["kl_loss"]) and torch.isfinite(losses["aux_loss"])
    xt2 = core.engine.q_sample(x0.cuda(), t.cuda(), seed=11)
    r = core.engine.vb_terms(x0.cuda(), xt2, t.cuda())
    mask = (t == 0).float().cuda()
    assert torch.allclose(losses["kl_loss"], ((mask * r["decoder_nll"] + (1 - mask) * r["kl"]) / pt.cuda()).mean())
    outputs2, losses2 = core(x0.cuda(), is_train=False)          # sample_time / q_sample from torch's generator
    assert "aux_loss" not in losses2 and int(core.Lt_count.sum()) == 2 * B
