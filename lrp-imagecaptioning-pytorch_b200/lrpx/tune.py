"""lrp_tune training step (reference train.py:211-233, ``train_lrp``), data-parallel over GPUs.

    predictions, weighted_predictions, L = model.forwardlrp_context(imgs, caps, caplens, rev_word_map)
    loss = CE(predictions, caps[:, 1:L+1]) + CE(weighted_predictions, caps[:, 1:L+1])       (ignore <pad>)
    optimizer.zero_grad(); loss.backward(); clip_gradient(optimizer, grad_clip); optimizer.step()

The LRP weights inside ``forwardlrp_context`` are per-sample constants (``lrpx_fc_lrp_weights_f32``, computed under
no_grad), so sharding the batch across ranks needs no exchange on the forward path; the only collective is the
gradient all-reduce, done by torch's DistributedDataParallel (NCCL over NVLink on the GPU box, gloo in the CPU
tests), bucketed and overlapped with the backward pass.  Each rank feeds its own slice of the global batch.
"""
import torch
import torch.distributed as dist
import torch.nn as nn


def clip_gradient(optimizer, grad_clip):
    """models/modelutils.py:102-112 — element-wise clamp of every gradient."""
    for group in optimizer.param_groups:
        for param in group['params']:
            if param.grad is not None:
                param.grad.data.clamp_(-grad_clip, grad_clip)


class _TuneForward(nn.Module):
    """``forward`` = ``model.forwardlrp_context`` so that DistributedDataParallel sees the call."""

    def __init__(self, model, rev_word_map):
        super().__init__()
        self.model = model
        self.rev_word_map = rev_word_map

    def forward(self, imgs, caps, caplens):
        pred, wpred, max_length = self.model.forwardlrp_context(imgs, caps, caplens, self.rev_word_map)
        return pred, wpred, torch.tensor(max_length)


class LrpTuneStep:
    """One ``train_lrp`` iteration per call.  ``model`` is any module with the reference's
    ``forwardlrp_context(imgs, caps, caplens, rev_word_map) -> (predictions, weighted_predictions, max_length)``
    (GridTDModel / AOAModel and their BU twins).  With an initialised process group the gradients are averaged
    over the ranks by DistributedDataParallel."""

    def __init__(self, model, word_map, optimizer=None, lr=1e-4, grad_clip=None, fix_encoder=True, tc_encoder="fp32"):
        """``tc_encoder``: with a fixed CNN (the reference's default, train.py:100-104) the VGG encoder forward runs on
        the tcgen05 engine ('fp32' = bf16x3, features within ~1.5e-4 of the library's fp32 convolutions; 'bf16'; None =
        the torch / cuDNN forward).  Ignored for models without ``use_tc_encoder`` and while the encoder is trained."""
        self.model = model
        self.word_map = word_map
        self.rev_word_map = {v: k for k, v in word_map.items()}
        if fix_encoder:                                     # train.py:100-104 ("Training with fixed CNN")
            for name, p in model.named_parameters():
                if 'img_encoder' in name:
                    p.requires_grad = False
        if fix_encoder and tc_encoder and hasattr(model, "use_tc_encoder"):
            model.use_tc_encoder(tc_encoder)
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = optimizer or torch.optim.Adam(params=params, lr=lr, betas=(0.8, 0.999))     # train.py:107-109
        self.grad_clip = grad_clip
        self.criterion = nn.CrossEntropyLoss(ignore_index=word_map['<pad>'])                        # train.py:124
        fwd = _TuneForward(model, self.rev_word_map)
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.distributed:
            dev = next(model.parameters()).device
            ids = [dev.index] if dev.type == "cuda" else None
            fwd = nn.parallel.DistributedDataParallel(fwd, device_ids=ids, broadcast_buffers=False)
        self.fwd = fwd

    @staticmethod
    def shard(batch, rank, world):
        """This rank's contiguous slice of a global batch (tensors with the batch on dim 0).  Balanced like
        ``shard.image_range`` (slices differ by at most one sample); every rank must get at least one sample — an
        empty slice would leave that rank out of DistributedDataParallel's gradient all-reduce and hang the others."""
        n = batch[0].shape[0]
        if n < world:
            raise ValueError(f"LrpTuneStep.shard: a batch of {n} samples cannot feed {world} ranks "
                             "(drop or pad the last short batch of the epoch)")
        base, extra = divmod(n, world)
        lo = rank * base + min(rank, extra)
        hi = lo + base + (1 if rank < extra else 0)
        return tuple(t[lo:hi] for t in batch)

    def losses(self, imgs, caps, caplens):
        self.model.train()
        pred, wpred, max_length = self.fwd(imgs, caps, caplens)
        max_length = int(max_length)
        targets = caps[:, 1:max_length + 1].contiguous().view(-1)
        if not self.distributed:
            loss_standard = self.criterion(pred.contiguous().view(-1, pred.size(2)), targets)
            loss_lrp = self.criterion(wpred.contiguous().view(-1, wpred.size(2)), targets)
            return loss_lrp + loss_standard, loss_standard, loss_lrp
        # The reference's loss is the mean over the non-<pad> tokens of the WHOLE batch (train.py:223-226).  Ranks
        # hold different token counts (ragged captions, slices differing by one sample), so the plain average of
        # per-rank means DistributedDataParallel would produce is a different number: each rank contributes
        # world * (sum of its token losses) / (global token count), whose rank average is the global token mean.
        pad = self.criterion.ignore_index
        n_tok = (targets != pad).sum().to(torch.float32)
        total = n_tok.detach().clone()
        dist.all_reduce(total)                              # one scalar; the gradient all-reduce stays DDP's
        scale = dist.get_world_size() / total.clamp(min=1.0)
        ce = nn.functional.cross_entropy
        loss_standard = ce(pred.contiguous().view(-1, pred.size(2)), targets, ignore_index=pad, reduction="sum") * scale
        loss_lrp = ce(wpred.contiguous().view(-1, wpred.size(2)), targets, ignore_index=pad, reduction="sum") * scale
        return loss_lrp + loss_standard, loss_standard, loss_lrp

    def step(self, imgs, caps, caplens):
        """-> (loss, loss_standard, loss_lrp) as detached tensors (no host sync here)."""
        loss, ls, ll = self.losses(imgs, caps, caplens)
        self.optimizer.zero_grad()
        loss.backward()                                     # DDP all-reduces the gradient buckets here
        if self.grad_clip:
            clip_gradient(self.optimizer, self.grad_clip)
        self.optimizer.step()
        return loss.detach(), ls.detach(), ll.detach()


class RewardCriterion(nn.Module):
    """models/modelutils.py:31-46 — policy-gradient loss of the self-critical tuner: -logprob * reward, masked to the
    sampled words up to and including the first <pad>/<end> position (mask shifted by one step)."""

    def forward(self, input, seq, reward):
        input = input.contiguous().view(-1)
        reward = reward.contiguous().view(-1)
        mask = (seq.detach() > 0).float()
        mask = torch.cat([mask.new_ones(mask.size(0), 1), mask[:, :-1]], 1).contiguous().view(-1)
        return torch.sum(-input * reward * mask) / torch.sum(mask)


class LrpCiderTuneStep:
    """One ``trainciderlrp`` iteration per call (reference train.py:252-272): greedy baseline with ``model.sample``,
    sampled captions with ``model.sample_lrp`` (the LRP weights of every step come from ``lrpx_fc_lrp_weights_f32``),
    self-critical reward, RewardCriterion, gradient clamp, optimizer step.

    ``reward_fn(greedy_seq, all_caps, sampled_seq, word_map) -> (B, L) array / tensor``; None = the reference's
    ``get_self_critical_reward`` with its training weights (train.py:264: CIDEr 1, BLEU 0), restated in ``lrpx.scst``
    (host n-gram counting like the reference's pycocoevalcap scorers; pinned by fixture scst_reward).  Single process,
    like the reference's loop."""

    def __init__(self, model, word_map, reward_fn=None, optimizer=None, lr=1e-5, grad_clip=None, fix_encoder=True):
        self.model = model
        self.word_map = word_map
        self.rev_word_map = {v: k for k, v in word_map.items()}
        if reward_fn is None:
            from . import scst
            reward_fn = lambda greedy, caps, gen, wm: scst.self_critical_reward(greedy, caps, gen, wm, 1.0, 0.0)
        self.reward_fn = reward_fn
        if fix_encoder:
            for name, p in model.named_parameters():
                if 'img_encoder' in name:
                    p.requires_grad = False
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = optimizer or torch.optim.Adam(params=params, lr=lr, betas=(0.8, 0.999))
        self.grad_clip = grad_clip
        self.criterion = RewardCriterion()

    def step(self, imgs, all_caps, caplens):
        """-> (loss, mean reward) as detached tensors."""
        m = self.model
        m.eval()
        with torch.no_grad():
            greedy_res, _, _ = m.sample(imgs, self.word_map, caplens)                                  # train.py:259-261
        m.train()
        gen_result, sample_logprobs, _ = m.sample_lrp(imgs, self.rev_word_map, self.word_map, caplens,
                                                      opt={'sample_method': 'sample'})                 # :263
        reward = torch.as_tensor(self.reward_fn(greedy_res, all_caps, gen_result, self.word_map),
                                 dtype=torch.float32, device=sample_logprobs.device)                   # :264-265
        loss = self.criterion(sample_logprobs, gen_result.data, reward)
        self.optimizer.zero_grad()
        loss.backward()
        if self.grad_clip:
            clip_gradient(self.optimizer, self.grad_clip)
        self.optimizer.step()
        return loss.detach(), reward[:, 0].mean().detach()
