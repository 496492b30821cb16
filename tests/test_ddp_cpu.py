"""world_size-2 gloo test (CPU) of the data-parallel host logic: flat gradient bucket, zero-fill of
None gradients, averaging, parameter broadcast (lanedetection_end2end_b200/ddp.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lanedetection_end2end_b200.ddp import FlatGradAllReduce, broadcast_parameters
    torch.manual_seed(100 + rank)                     # replicas start different ...
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2), torch.nn.Linear(2, 2))
    broadcast_parameters(net, 0)                      # ... and are made identical
    p0 = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    x = torch.full((5, 4), float(rank + 1))
    net[1](net[0](x)).sum().backward()                # net[2] never used -> its grads stay None
    red = FlatGradAllReduce(net)
    local = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
    flat = red()
    # numpy arrays are pickled by value (torch tensors would be passed as shared-memory fds, which die
    # with this process)
    after = [None if p.grad is None else p.grad.numpy().copy() for p in net.parameters()]
    flat1 = flat.numpy().copy()
    # second step with the gradients ATTACHED to the flat buffer (what GraphedTrainStep does): backward accumulates
    # straight into the buffer, the reducer only all-reduces
    red.attach()
    torch._foreach_zero_([p.grad for p in net.parameters() if p.grad is not None])
    net[1](net[0](x * 2)).sum().backward()
    assert all(p.grad is None or p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
    local2 = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
    red()
    q.put((rank, p0.numpy(), [None if g is None else g.numpy() for g in local], flat1, after,
           [None if g is None else g.numpy() for g in local2],
           [None if p.grad is None else p.grad.numpy().copy() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    t = lambda v: None if v is None else torch.from_numpy(v)
    res = [(r, t(p0), [t(x) for x in l], t(f), [t(x) for x in g], [t(x) for x in l2], [t(x) for x in g2])
           for r, p0, l, f, g, l2, g2 in res]
    (_, p0a, la, fa, ga, l2a, g2a), (_, p0b, lb, fb, gb, l2b, g2b) = res
    for a2, b2, g2 in zip(l2a, l2b, g2a):             # attached step: same averaging, no pack / unpack
        if a2 is not None:
            torch.testing.assert_close(g2, (a2 + b2) / 2)
    assert torch.equal(p0a, p0b)                      # broadcast worked
    assert torch.equal(fa, fb)                        # every rank holds the same reduced buffer
    off = 0
    for a, b, g in zip(la, lb, ga):
        n = (a if a is not None else b if b is not None else torch.zeros(0)).numel()
        if a is None:
            assert g is None                          # stays None on the module ...
            continue
        want = (a + b) / 2
        torch.testing.assert_close(g, want)
        torch.testing.assert_close(fa[off:off + n].view_as(want), want)
        off += n
    # ... but the unused layer contributed zeros to the flat buffer (fixed layout on every rank)
    assert float(fa[off:].abs().sum()) == 0.0 and fa.numel() - off == 2 * 2 + 2


def _worker_valid(rank, world, port, q):
    import numpy as np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lanedetection_end2end_b200.ddp import lane_valid_scale
    from lanedetection_end2end_b200.Loss_crit import backprojection_loss
    from lanedetection_end2end_b200.Networks.utils import define_args
    args = define_args().parse_args(["--image_dir", "x", "--gt_dir", "y", "--nclasses", "4", "--order", "3", "--no_cuda"])
    crit = backprojection_loss(args)
    beta, xgt, valid = _valid_shard(rank)
    betas = [beta[:, l].unsqueeze(-1).clone().requires_grad_(True) for l in range(4)]
    scale = lane_valid_scale(valid, 4)
    loss, _ = crit.forward_lanes(betas, xgt, valid, lane_scale=scale)
    loss.backward()
    g = torch.stack([b.grad.squeeze(-1) for b in betas], 1)          # [B, L, n]
    lt = loss.detach().clone().reshape(1)
    dist.all_reduce(lt)                                                # what DDP does: mean over ranks
    q.put((rank, float(lt) / world, (g / world).numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _valid_shard(rank):
    """Shards with DIFFERENT numbers of valid samples per lane (lane 3 of rank 1 has none at all)."""
    g = torch.Generator().manual_seed(7 + rank)
    B = 3
    beta = torch.randn(B, 4, 4, generator=g, dtype=torch.float64) * torch.tensor([1e-6, 1e-3, 0.5, 200.0], dtype=torch.float64)
    xgt = torch.rand(B, 4, 56, generator=g, dtype=torch.float64) * 500
    valid = (torch.rand(B, 4, 56, generator=g) > (0.3 + 0.4 * rank)).double()
    if rank == 1:
        valid[:, 3] = 0
    return beta, xgt, valid


def test_batch_global_valid_normaliser_two_ranks():
    """ddp.lane_valid_scale: with it, the rank-mean of the per-shard losses and gradients equals the single-process value on
    the concatenated batch (the reference's batch-global sum(valid), BP/Loss_crit.py:215)."""
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_valid, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        r, loss, g = q.get(timeout=120)
        got[r] = (loss, g)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from lanedetection_end2end_b200.Loss_crit import backprojection_loss
    from lanedetection_end2end_b200.Networks.utils import define_args
    args = define_args().parse_args(["--image_dir", "x", "--gt_dir", "y", "--nclasses", "4", "--order", "3", "--no_cuda"])
    crit = backprojection_loss(args)
    shards = [_valid_shard(r) for r in range(2)]
    beta = torch.cat([s[0] for s in shards]).requires_grad_(True)
    xgt, valid = torch.cat([s[1] for s in shards]), torch.cat([s[2] for s in shards])
    # the reference's own per-lane calls on the whole batch (BP/main.py:297-305)
    loss = sum(crit(beta[:, l].unsqueeze(-1), xgt[:, l], valid[:, l])[0] for l in range(4)) / 4
    loss.backward()
    assert abs(got[0][0] - float(loss)) <= 1e-12 * abs(float(loss)) and abs(got[1][0] - float(loss)) <= 1e-12 * abs(float(loss))
    want = beta.grad.numpy()
    have = np.concatenate([got[0][1], got[1][1]])       # each rank holds the (already 1/world-averaged) gradient of ITS shard
    np.testing.assert_allclose(have, want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
