"""CPU, build container only (skipped when /root/reference is absent): pieces of the oracle against the UNMODIFIED
reference beyond what tests/golden/make_golden.py already asserts while generating the fixtures."""
import numpy as np
import pytest
import torch

import ref_harness as rh
from oracle import layoutdm_oracle as O

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    vocab, spec = O.RICO25, O.ModelSpec(layers=1)
    sd = O.make_weights(vocab, O.ModelSpec(), seed=1)
    model, tok = rh.build_reference("rico25", state_dict=sd)
    return model, tok, vocab


def test_q_sample_ids_matches_reference_q_sample(ref):
    """forward (corruption) process: oracle.q_sample_ids == reference q_sample per attribute (constrained.py:223-230),
    with the Gumbel uniforms injected through torch.rand_like"""
    model, tok, vocab = ref
    core = model.model.module
    B, S, C, T = 6, vocab.S, vocab.C, 100
    g = torch.Generator().manual_seed(3)
    x0 = torch.empty(B, S, dtype=torch.long)
    for a in range(5):
        ids = torch.tensor(vocab.group_full_ids(a)[:-1])               # normal classes + PAD (no MASK in x0)
        x0[:, a::5] = ids[torch.randint(0, len(ids), (B, 25), generator=g)]
    t = torch.tensor([0, 1, 37, 64, 98, 99])
    u = O.uniforms(77, 0, 2, 0, B, S, C)
    want = O.q_sample_ids(x0, t, T, vocab, O.group_schedules(T, vocab), u)
    orig = torch.rand_like
    got = torch.empty_like(x0)
    try:
        for a, key in enumerate(tok.var_names):
            idx = torch.tensor(vocab.group_full_ids(a))
            K = len(idx)
            part = core.converter.f_to_p_id(x0[:, a::5], key)
            log_x0 = torch.log(torch.nn.functional.one_hot(part, K).permute(0, 2, 1).float().clamp(min=1e-30))
            ua = torch.from_numpy(u)[:, a::5][..., idx].permute(0, 2, 1).contiguous()      # (B, K, 25)
            torch.rand_like = lambda x, **kw: ua
            log_xt = core.q_sample(log_x_start=log_x0, t=t, key=key)
            got[:, a::5] = core.converter.p_to_f_id(log_xt.argmax(1), key)
    finally:
        torch.rand_like = orig
    assert torch.equal(got, want)
    # sanity: late timesteps are mostly MASK, early ones mostly unchanged
    assert (want[5] == vocab.mask_id).float().mean() > 0.9 and (want[0] == x0[0]).float().mean() > 0.9


def test_decode_matches_reference_tokenizer(ref):
    model, tok, vocab = ref
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, vocab.C, (32, vocab.S), generator=g)
    a, b = tok.decode(ids.clone()), O.decode_ids(ids, vocab)
    for k in ("bbox", "label", "mask"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("cond_type", ["c", "cwh", "gt", "refinement"])
def test_make_cond_matches_reference_get_cond(ref, cond_type):
    """cond construction: oracle.make_cond == the reference's get_cond (task.py:27-151) on a fake PyG batch, incl.
    boxes outside [0, 1] and on the rounding boundaries of the linear quantisation"""
    model, tok, vocab = ref
    rh._setup_path()
    from trainer.data.util import sparse_to_dense
    from trainer.helpers.task import get_cond
    batch = rh.synthetic_layouts(48, vocab.n_cat, seed=3)
    batch.x = batch.x * 1.3 - 0.15                              # some coordinates below 0 / above 1
    batch.x[::7] = (torch.arange(batch.x[::7].numel()).view(-1, 4) % 33).float() / 32.0 + 1.0 / 64.0   # exact .5 bin boundaries
    bbox, label, _, mask = sparse_to_dense(batch)
    torch.manual_seed(11)
    want = get_cond(batch, tok, cond_type=cond_type, model_type="LayoutDM")
    if cond_type == "refinement":
        torch.manual_seed(11)
        bbox = bbox + torch.normal(0, std=0.1, size=bbox.size())   # the draw of task.py:127
    got = O.make_cond(label, bbox, mask, vocab, cond_type)
    for k in ("seq", "mask") + (("seq_orig",) if cond_type == "refinement" else ()):
        assert torch.equal(want[k], got[k]), k
    if cond_type != "gt":
        assert torch.equal(want["num_element"], got["num_element"])
