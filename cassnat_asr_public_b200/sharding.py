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


class GradientBuckets:
    """Data-parallel gradient exchange of the training step (the reference wraps the model in
    DistributedDataParallel, src/tasks/base_task.py + cassnat_task.py:291): parameters are grouped, in
    reverse registration order (roughly the order backward produces them), into flat buckets; a
    post-accumulate hook launches the bucket's all_reduce asynchronously as soon as its last gradient
    exists, so the exchange overlaps the rest of backward on NCCL's own stream.  `finish()` waits,
    averages and scatters the result back into `p.grad`.  comm_dtype=torch.bfloat16 halves the bytes on
    NVLink (fp32 master gradients are restored on the way back)."""

    def __init__(self, params, group=None, bucket_bytes=32 << 20, comm_dtype=torch.float32, hooks=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.comm_dtype = comm_dtype
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        esize = torch.empty(0, dtype=comm_dtype).element_size()
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * esize
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._flat = [None] * len(self.buckets)
        self._pending = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        # hooks=False: no overlap; finish() launches every bucket itself (used when backward is a replayed
        # CUDA graph, where no autograd hook fires)
        self._hooks = [p.register_post_accumulate_grad_hook(self._ready) for p in self.params] \
            if (self.world > 1 and hooks) else []
        self.reset()

    def reset(self):
        """Call before every backward whose gradients are to be exchanged (finish() does it for the next step).  With
        gradient accumulation (cassnat_task.py:290-292) call `reset()` + `finish()` around the LAST micro-batch only
        and keep `hooks_enabled = False` for the earlier ones, the equivalent of DDP's no_sync()."""
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    hooks_enabled = True

    def _ready(self, p):
        if not self.hooks_enabled:
            return
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0 and self._work[i] is None:
            self._launch(i)

    def _launch(self, i):
        b = self.buckets[i]
        n = sum(p.numel() for p in b)
        if self._flat[i] is None or self._flat[i].device != b[0].grad.device:
            self._flat[i] = torch.empty(n, dtype=self.comm_dtype, device=b[0].grad.device)
        flat, o = self._flat[i], 0
        for p in b:
            flat[o:o + p.numel()].copy_(p.grad.reshape(-1))
            o += p.numel()
        self._work[i] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Call after loss.backward(): waits for every bucket and writes the averaged gradients back."""
        if self.world == 1:
            return
        for i, b in enumerate(self.buckets):
            if self._work[i] is None:              # parameters that received no gradient this step: every rank still
                for p in b:                        # launches the bucket (zeros), so the collectives always match up
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(i)
            self._work[i].wait()
            flat, o = self._flat[i], 0
            for p in b:
                p.grad.copy_(flat[o:o + p.numel()].reshape(p.shape)).div_(self.world)
                o += p.numel()
        self.reset()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


class FlatGradients:
    """Gradients of all parameters as views of ONE flat fp32 buffer, so that the data-parallel exchange is a
    single in-place all_reduce with no gather/scatter copies.  Meant for the graph-replayed training step:
    the buffer is static memory, backward accumulates into the views (`p.grad` stays defined), `zero()` and
    the backward are captured, `all_reduce()` runs after the replay."""

    def __init__(self, params, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        self.attach()

    def attach(self):
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view(p.shape)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world)


class OverlappedFlatGradients(FlatGradients):
    """FlatGradients whose exchange overlaps backward: the parameters are laid out in reverse registration order (the
    order backward produces their gradients) and cut into contiguous buckets of the flat buffer; a post-accumulate
    hook launches a bucket's all_reduce (NCCL's own stream, async) as soon as its last gradient is complete, while
    the rest of backward keeps running.  comm_dtype=torch.bfloat16 sends bf16 on NVLink (half the bytes; fp32 master
    gradients restored, averaged, on the way back).  Everything -- hooks, casts, collectives, the final wait -- is
    stream work, so a whole step (zero(), forward, backward, finish()) can be captured in ONE CUDA graph and replayed:
    the all_reduces then overlap the replayed backward as well.
    Reference: DistributedDataParallel's bucketed gradient all_reduce (src/tasks/base_task.py:82-83)."""

    def __init__(self, params, group=None, bucket_bytes=16 << 20, comm_dtype=torch.bfloat16):
        ordered = [p for p in reversed(list(params)) if p.requires_grad]
        super().__init__(ordered, group)
        self.comm_dtype = comm_dtype
        self.bounds, lo, o = [], 0, 0                       # element ranges [lo, hi) of the buckets in the flat buffer
        esize = torch.empty(0, dtype=comm_dtype).element_size()
        self._bucket_of = {}
        for p in self.params:
            self._bucket_of[id(p)] = len(self.bounds)
            o += p.numel()
            if (o - lo) * esize >= bucket_bytes:
                self.bounds.append((lo, o))
                lo = o
        if o > lo:
            self.bounds.append((lo, o))
        self._count = [0] * len(self.bounds)
        for p in self.params:
            self._count[self._bucket_of[id(p)]] += 1
        self.comm = None if comm_dtype == torch.float32 else torch.empty(self.flat.numel(), dtype=comm_dtype, device=self.flat.device)
        self._hooks = [p.register_post_accumulate_grad_hook(self._ready) for p in self.params] if self.world > 1 else []
        self.hooks_enabled = True
        self.begin()

    def begin(self):
        """Before every backward that is to be exchanged (zero() calls it)."""
        self._pending = list(self._count)
        self._work = [None] * len(self.bounds)

    def zero(self):
        self.flat.zero_()
        self.begin()

    def _ready(self, p):
        if not self.hooks_enabled:
            return
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0 and self._work[i] is None:
            self._launch(i)

    def _launch(self, i):
        lo, hi = self.bounds[i]
        buf = self.flat[lo:hi]
        if self.comm is not None:
            buf = self.comm[lo:hi]
            buf.copy_(self.flat[lo:hi])
        self._work[i] = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """After backward: launches what no hook launched (same decision on every rank), waits, averages."""
        if self.world == 1:
            return
        for i, (lo, hi) in enumerate(self.bounds):
            if self._work[i] is None:
                self._launch(i)
        inv = 1.0 / self.world
        for i, (lo, hi) in enumerate(self.bounds):
            self._work[i].wait()
            if self.comm is not None:
                # cast back and average in one pass (the product is formed in the comm dtype: exact for power-of-two
                # world sizes, otherwise one more rounding of the size the bf16 exchange already has)
                torch.mul(self.comm[lo:hi], inv, out=self.flat[lo:hi])
            else:
                self.flat[lo:hi].mul_(inv)

    all_reduce = finish

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
