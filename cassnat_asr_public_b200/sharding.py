"""Utterance sharding across one process per GPU, and the hypothesis / error-count gather.

The reference decodes multi-GPU by splitting the utterance list into `nj` shards, one process per
device, and concatenating the result files (egs/librispeech/run_hubert.sh:94-116); training uses
DDP plus one 6-float all_reduce (src/tasks/cassnat_task.py:186-189).  Utterances are independent on
the whole alignment-conditioned path, so there is no data-path collective: each rank decodes a
contiguous block and the only exchange is this gather of padded token ids (+ lengths) and the
(errors, tokens) all_reduce for WER.  Works on any torch.distributed backend (nccl on the box,
gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world_size):
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world_size ranks get one more."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_hypotheses(hyps, pad_len, device):
    """list of token-id lists -> (ids int32 [n, pad_len] zero padded, lens int32 [n])."""
    import numpy as np
    n = len(hyps)
    ids = np.zeros((n, pad_len), dtype=np.int32)
    lens = np.zeros(n, dtype=np.int32)
    for i, h in enumerate(hyps):
        if len(h) > pad_len:
            raise ValueError(f"hypothesis of length {len(h)} exceeds pad_len {pad_len}")
        ids[i, :len(h)] = h
        lens[i] = len(h)
    return torch.from_numpy(ids).to(device), torch.from_numpy(lens).to(device)


def unpack_hypotheses(ids, lens):
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    return [ids[i, :int(lens[i])].tolist() for i in range(ids.shape[0])]


def gather_hypotheses(hyps, n_total, pad_len, device, group=None):
    """All ranks call this with the hypotheses of their own contiguous shard (shard_bounds order);
    every rank gets the n_total hypotheses in global utterance order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return list(hyps)
    rank = dist.get_rank(group)
    per_rank = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
    ids, lens = pack_hypotheses(hyps, pad_len, device)
    packed = torch.cat([ids, lens.unsqueeze(1)], 1)            # lengths ride in the last column: one collective
    if packed.size(0) < per_rank:                                # equalise shapes for all_gather
        packed = torch.cat([packed, packed.new_zeros(per_rank - packed.size(0), pad_len + 1)], 0)
    everything = torch.empty(world * per_rank, pad_len + 1, dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(everything, packed.contiguous(), group=group)
    everything = everything.cpu().numpy()
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        blk = everything[r * per_rank:r * per_rank + (hi - lo)]
        out.extend(blk[i, :int(blk[i, pad_len])].tolist() for i in range(hi - lo))
    return out


def reduce_error_counts(errors, tokens, device, group=None):
    """Sum of (errors, tokens) over ranks -> python ints (the WER numerator / denominator)."""
    t = torch.tensor([float(errors), float(tokens)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0].item()), int(t[1].item())


def edit_distance(a, b):
    """Levenshtein distance between two token lists (the reference uses the `editdistance`
    C extension in src/utils/wer.py:7-52; this is the same metric)."""
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]
