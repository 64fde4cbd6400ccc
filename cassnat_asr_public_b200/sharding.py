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
    n = len(hyps)
    ids = torch.zeros(n, pad_len, dtype=torch.int32)
    lens = torch.zeros(n, dtype=torch.int32)
    for i, h in enumerate(hyps):
        if len(h) > pad_len:
            raise ValueError(f"hypothesis of length {len(h)} exceeds pad_len {pad_len}")
        ids[i, :len(h)] = torch.tensor(h, dtype=torch.int32)
        lens[i] = len(h)
    return ids.to(device), lens.to(device)


def unpack_hypotheses(ids, lens):
    ids, lens = ids.cpu(), lens.cpu()
    return [ids[i, :int(lens[i])].tolist() for i in range(ids.size(0))]


def gather_hypotheses(hyps, n_total, pad_len, device, group=None):
    """All ranks call this with the hypotheses of their own contiguous shard (shard_bounds order);
    every rank gets the n_total hypotheses in global utterance order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return list(hyps)
    rank = dist.get_rank(group)
    per_rank = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
    ids, lens = pack_hypotheses(hyps, pad_len, device)
    if ids.size(0) < per_rank:                      # equalise shapes for all_gather
        ids = torch.cat([ids, ids.new_zeros(per_rank - ids.size(0), pad_len)], 0)
        lens = torch.cat([lens, lens.new_zeros(per_rank - lens.size(0))], 0)
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(all_ids, ids, group=group)
    dist.all_gather(all_lens, lens, group=group)
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        out.extend(unpack_hypotheses(all_ids[r][:hi - lo], all_lens[r][:hi - lo]))
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
