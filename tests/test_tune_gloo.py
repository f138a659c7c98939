"""CPU, world_size 2 over gloo: host logic of the lrp_tune step (lrpx.tune.LrpTuneStep, reference train.py:211-233):
loss assembly, gradient clamp, and that two ranks each stepping on half of a batch end up with the same parameters as
one process stepping on the whole batch (DistributedDataParallel gradient averaging; equal caption lengths so the
per-rank token means equal the global mean)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

import synth
from lrpx import tune


class _StubCaptioner(nn.Module):
    """Pure-torch stand-in with the reference's forwardlrp_context contract (the real models call the CUDA kernel
    lrpx_fc_lrp_weights_f32 inside, which has no CPU form).  The "LRP weights" are constants under no_grad."""

    def __init__(self, V, E=8, H=12):
        super().__init__()
        self.embedding = nn.Embedding(V, E)
        self.img = nn.Linear(6, H)
        self.rnn = nn.Linear(E + H, H)
        self.fc = nn.Linear(H, V)

    def forwardlrp_context(self, imgs, caps, caplens, rev_word_map):
        max_length = int(max(caplens)) - 1
        h = torch.tanh(self.img(imgs))
        preds, wpreds = [], []
        for t in range(max_length):
            h = torch.tanh(self.rnn(torch.cat((self.embedding(caps[:, t]), h), -1)))
            with torch.no_grad():
                w = h.detach().abs() / h.detach().abs().amax(1, keepdim=True) + 1.0
            preds.append(self.fc(h))
            wpreds.append(self.fc(h * w))
        return torch.stack(preds, 1), torch.stack(wpreds, 1), max_length


def _data(V, B=8, L=5, ragged=False):
    g = torch.Generator().manual_seed(3)
    imgs = torch.randn(B, 6, generator=g)
    caps = torch.randint(1, V - 4, (B, L), generator=g)
    caps[:, 0] = V - 2
    caplens = torch.full((B,), L)
    if ragged:                                  # caption b has 2 + (b % (L-1)) tokens, <pad> = 0 after them
        caplens = torch.tensor([2 + (b % (L - 1)) for b in range(B)])
        for b in range(B):
            caps[b, int(caplens[b]):] = 0
    return imgs, caps, caplens


def _single_process_reference(V, steps, **kw):
    torch.manual_seed(0)
    m = _StubCaptioner(V)
    st = tune.LrpTuneStep(m, synth.word_map(V), lr=1e-2, grad_clip=0.05)
    for _ in range(steps):
        loss, ls, ll = st.step(*_data(V, **kw))
    return [p.detach().clone() for p in m.parameters()], float(loss)


def _worker(rank, world, port, V, steps, ref_params, kw=None):
    kw = kw or {}
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = _StubCaptioner(V)
        st = tune.LrpTuneStep(m, synth.word_map(V), lr=1e-2, grad_clip=0.05)
        assert st.distributed
        for _ in range(steps):
            st.step(*tune.LrpTuneStep.shard(_data(V, **kw), rank, world))
        for p, r in zip(m.parameters(), ref_params):
            assert torch.allclose(p.detach(), r, rtol=1e-5, atol=1e-6), float((p.detach() - r).abs().max())
    finally:
        dist.destroy_process_group()


def test_loss_and_clip_single_process():
    V = 30
    torch.manual_seed(0)
    m = _StubCaptioner(V)
    st = tune.LrpTuneStep(m, synth.word_map(V), lr=1e-2, grad_clip=0.05)
    imgs, caps, caplens = _data(V)
    loss, ls, ll = st.losses(imgs, caps, caplens)
    pred, wpred, L = m.forwardlrp_context(imgs, caps, caplens, None)
    ce = nn.CrossEntropyLoss(ignore_index=0)
    tgt = caps[:, 1:L + 1].reshape(-1)
    assert torch.allclose(ls, ce(pred.reshape(-1, V), tgt)) and torch.allclose(ll, ce(wpred.reshape(-1, V), tgt))
    assert torch.allclose(loss, ls + ll)
    before = [p.detach().clone() for p in m.parameters()]
    st.step(imgs, caps, caplens)
    assert all(p.grad is None or float(p.grad.abs().max()) <= 0.05 + 1e-9 for p in m.parameters())
    assert any(not torch.equal(b, p.detach()) for b, p in zip(before, m.parameters()))


def test_two_ranks_equal_single_process_gloo():
    V, steps = 30, 3
    ref_params, _ = _single_process_reference(V, steps)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, V, steps, ref_params), nprocs=2, join=True)


def test_two_ranks_uneven_ragged_batch_equal_single_process_gloo():
    """5 samples over 2 ranks (3 + 2) with ragged captions: the token-weighted loss makes the averaged gradients equal
    to the single-process step on the whole batch (the reference's mean over all non-<pad> tokens)."""
    V, steps = 30, 3
    kw = dict(B=5, ragged=True)
    ref_params, _ = _single_process_reference(V, steps, **kw)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, V, steps, ref_params, kw), nprocs=2, join=True)


def test_shard_is_balanced_and_refuses_empty_ranks():
    import pytest
    b = (torch.arange(5),)
    sizes = [tune.LrpTuneStep.shard(b, r, 4)[0].shape[0] for r in range(4)]
    assert sizes == [2, 1, 1, 1]
    assert torch.equal(torch.cat([tune.LrpTuneStep.shard(b, r, 4)[0] for r in range(4)]), b[0])
    with pytest.raises(ValueError):
        tune.LrpTuneStep.shard((torch.arange(3),), 0, 4)


class _StubSampler(nn.Module):
    """Pure-torch stand-in with the reference's sample / sample_lrp contract (train.py:259-263)."""

    def __init__(self, V, H=10):
        super().__init__()
        self.img = nn.Linear(6, H)
        self.fc = nn.Linear(H, V)

    def _roll(self, imgs, L, greedy):
        h = torch.tanh(self.img(imgs))
        logp = torch.log_softmax(self.fc(h), -1)
        seq, lps = [], []
        for t in range(L):
            it = logp.argmax(-1) if greedy else torch.multinomial(logp.exp(), 1)[:, 0]
            seq.append(it)
            lps.append(logp.gather(1, it[:, None])[:, 0])
        return torch.stack(seq, 1), torch.stack(lps, 1), L

    def sample(self, imgs, word_map, caplens, opt={}):
        return self._roll(imgs, int(max(caplens)) - 1, True)

    def sample_lrp(self, imgs, rev_word_map, word_map, caplens, opt={}):
        assert opt.get('sample_method') == 'sample'
        return self._roll(imgs, int(max(caplens)) - 1, False)


def test_cider_tune_step_host_logic():
    """lrpx.tune.LrpCiderTuneStep (train.py:252-272): greedy baseline under no_grad / eval, sampled captions in train
    mode, reward handed to RewardCriterion, clamp, step; RewardCriterion's mask (modelutils.py:40-41)."""
    V = 30
    torch.manual_seed(0)
    m = _StubSampler(V)
    seen = {}

    def reward_fn(greedy, all_caps, sampled, word_map):
        seen["greedy_requires_grad"] = greedy.requires_grad
        seen["all_caps"] = all_caps
        return (sampled != greedy).float().numpy() - 0.25

    st = tune.LrpCiderTuneStep(m, synth.word_map(V), reward_fn, lr=1e-2, grad_clip=0.05)
    before = [p.detach().clone() for p in m.parameters()]
    imgs, caps, caplens = _data(V)
    loss, rew = st.step(imgs, "refs", caplens)
    assert seen["all_caps"] == "refs" and seen["greedy_requires_grad"] is False and m.training
    assert torch.isfinite(loss) and -0.25 <= float(rew) <= 0.75
    assert all(p.grad is None or float(p.grad.abs().max()) <= 0.05 + 1e-9 for p in m.parameters())
    assert any(not torch.equal(b, p.detach()) for b, p in zip(before, m.parameters()))
    crit = tune.RewardCriterion()
    logp = torch.tensor([[-1.0, -2.0, -3.0], [-0.5, -0.5, -0.5]])
    seq = torch.tensor([[5, 0, 0], [4, 4, 4]])
    reward = torch.tensor([[2.0, 2.0, 2.0], [1.0, 1.0, 1.0]])
    # mask = [1, seq>0 shifted right]: row 0 -> [1,1,0], row 1 -> [1,1,1]
    want = (1.0 * 2 + 2.0 * 2 + 0.5 + 0.5 + 0.5) / 5
    assert abs(float(crit(logp, seq, reward)) - want) < 1e-6
