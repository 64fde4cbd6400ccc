"""CPU: the multi-GPU host logic (utterance sharding, hypothesis gather, WER count reduction) on a
world_size-2 gloo group — the N > 1 path of bench.py without GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cassnat_asr_public_b200 import sharding


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 65, 1000):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip_and_edit_distance():
    hyps = [[1, 5, 6], [1], [1, 2, 3, 4, 5, 6, 7]]
    ids, lens = sharding.pack_hypotheses(hyps, 9, "cpu")
    assert sharding.unpack_hypotheses(ids, lens) == hyps
    assert sharding.edit_distance([1, 2, 3], [1, 3]) == 1 and sharding.edit_distance([], [4, 4]) == 2
    assert sharding.edit_distance([7, 8, 9], [7, 8, 9]) == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    mine = [[1] + [10 * i + k for k in range(i % 5)] for i in range(lo, hi)]       # hypothesis of utterance i
    everything = sharding.gather_hypotheses(mine, n_total, 8, "cpu")
    errs, toks = sharding.reduce_error_counts(errors=rank + 1, tokens=10 * (rank + 1), device="cpu")
    out_q.put((rank, everything, errs, toks))
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    world, n_total = 2, 7                       # uneven shards: 4 + 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[1] + [10 * i + k for k in range(i % 5)] for i in range(n_total)]
    for rank, everything, errs, toks in results:
        assert everything == want, rank
        assert (errs, toks) == (3, 30)


def _grad_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                      # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    for p in net[2].parameters():                             # a bucket member that never gets a gradient
        pass
    buckets = sharding.GradientBuckets(net.parameters(), bucket_bytes=64)     # tiny buckets -> several of them
    outs = []
    for step in range(2):                                     # buckets are reusable across steps
        net.zero_grad(set_to_none=True)
        x = torch.full((4, 6), float(rank + 1 + step))
        net(x).square().sum().backward()
        local = [p.grad.clone() for p in net.parameters()]
        buckets.finish()
        outs.append((local, [p.grad.clone() for p in net.parameters()]))
    out_q.put((rank, [([g.numpy() for g in l], [g.numpy() for g in a]) for l, a in outs]))
    dist.destroy_process_group()


def test_gradient_buckets_two_ranks_gloo():
    """Bucketed asynchronous gradient all_reduce == the mean of the per-rank gradients."""
    import numpy as np
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(2):
        local0, avg0 = results[0][step]
        local1, avg1 = results[1][step]
        for a, b, m0, m1 in zip(local0, local1, avg0, avg1):
            assert np.allclose(m0, (a + b) / 2, rtol=1e-6) and np.allclose(m1, m0)


def _flat_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    flat = sharding.FlatGradients(net.parameters())
    res = []
    for step in range(2):
        flat.zero()
        net(torch.full((3, 6), float(rank + 1 + step))).square().sum().backward()
        assert all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in net.parameters())     # still views of the buffer
        local = [p.grad.clone().numpy() for p in net.parameters()]
        flat.all_reduce()
        res.append((local, [p.grad.clone().numpy() for p in net.parameters()]))
    out_q.put((rank, res))
    dist.destroy_process_group()


def test_flat_gradients_two_ranks_gloo():
    import numpy as np
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flat_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(2):
        (l0, a0), (l1, a1) = results[0][step], results[1][step]
        for x, y, m0, m1 in zip(l0, l1, a0, a1):
            assert np.allclose(m0, (x + y) / 2, rtol=1e-6) and np.allclose(m0, m1)


def _overlap_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
    unused = torch.nn.Linear(3, 3)                                         # parameters no backward ever reaches
    params = list(net.parameters()) + list(unused.parameters())
    res = []

    def local_grads(inputs):
        """This rank's own gradients, computed without any exchange (buckets reduced during backward already hold
        the sum by the time backward returns, so p.grad cannot be read for that)."""
        gs = torch.autograd.grad(sum(net(x).square().sum() for x in inputs), list(net.parameters()))
        return [g.numpy() for g in gs] + [torch.zeros_like(p).numpy() for p in unused.parameters()]

    for comm in (torch.float32, torch.bfloat16):
        og = sharding.OverlappedFlatGradients(params, bucket_bytes=48, comm_dtype=comm)     # several tiny buckets
        assert len(og.bounds) >= 3
        for step in range(2):
            x = torch.full((3, 6), float(rank + 1 + step))
            local = local_grads([x])
            og.zero()
            net(x).square().sum().backward()
            launched = sum(w is not None for w in og._work)                # hooks fired during backward
            og.finish()
            res.append((str(comm), launched, local, [p.grad.clone().numpy() for p in params]))
        # gradient accumulation: first micro-batch without exchange (hooks off), second one exchanges the sum
        xs = [torch.full((3, 6), float(rank + 1)), torch.full((3, 6), float(rank + 2))]
        local = local_grads(xs)
        og.zero()
        og.hooks_enabled = False
        net(xs[0]).square().sum().backward()
        og.hooks_enabled = True
        og.begin()
        net(xs[1]).square().sum().backward()
        og.finish()
        res.append((str(comm) + "/accum", 0, local, [p.grad.clone().numpy() for p in params]))
        og.remove()
    out_q.put((rank, res))
    dist.destroy_process_group()


def test_overlapped_flat_gradients_two_ranks_gloo():
    """Bucketed all_reduce launched from post-accumulate hooks over slices of ONE flat buffer (fp32 and bf16 on the
    wire), incl. parameters that never receive a gradient and a gradient-accumulation step."""
    import numpy as np
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (tag, launched, l0, a0), (_, _, l1, a1) in zip(results[0], results[1]):
        tol = 1e-6 if "float32" in tag else 2e-2
        if "accum" not in tag:
            assert launched >= 2, tag                       # at least the early buckets went out during backward
        for x, y, m0, m1 in zip(l0, l1, a0, a1):
            assert np.allclose(m0, (x + y) / 2, rtol=tol, atol=tol * (np.abs(x).max() + 1e-6)), tag
            assert np.array_equal(m0, m1)
