"""Mirror of the self-critical-training pieces of the reference's models/modelutils.py:
``RewardCriterion`` (:30-45), ``array_to_str`` (:186-198), ``get_self_critical_reward`` (:200-238).  The n-gram scorers
live in ``lrpx.scst`` (own restatement of the reference's vendored pycocoevalcap CIDEr / BLEU).  The rest of that file
(checkpointing, GloVe loading, meters) is training plumbing outside the hot path."""
import torch
import torch.nn as nn

from lrpx import scst


class RewardCriterion(nn.Module):
    """reference :30-45: -(log p) * reward over the sampled tokens; a position counts while the PREVIOUS token was not
    <pad> (the mask is the token mask shifted right by one, first column 1)."""

    def forward(self, input, seq, reward):
        mask = (seq.detach() > 0).float()
        mask = torch.cat([mask.new_ones(mask.size(0), 1), mask[:, :-1]], 1).reshape(-1)
        output = -input.reshape(-1) * reward.reshape(-1) * mask
        return torch.sum(output) / torch.sum(mask)


def array_to_str(arr, rev_word_map, end_encode):
    """reference :186-198, quirk included: ``end_encode in out`` compares an id with words and never stops the loop."""
    out = []
    for i in range(len(arr)):
        if end_encode in out:
            break
        elif rev_word_map[int(arr[i])] not in ['<start>', '<pad>']:
            out.append(rev_word_map[int(arr[i])])
    return ' '.join(out)


def get_self_critical_reward(greedy_res, data_gts, gen_result, word_map, cider_reward_weight, bleu_reward_weight):
    """reference :200-238 -> numpy (batch, seq_length)."""
    return scst.self_critical_reward(greedy_res, data_gts, gen_result, word_map, cider_reward_weight, bleu_reward_weight)
