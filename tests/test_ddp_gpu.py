"""Multi-GPU correctness on hardware (SURVEY.md 4 iv / 8e): a 2-rank data-parallel step (one process per GPU, NCCL,
full replica, per-replica BatchNorm statistics, ONE flat-gradient all-reduce) must equal the mean of two 1-rank steps on
the same shards.  Skipped on a box with fewer than 2 GPUs (the driver's 1-GPU test tier)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _one_step(model, crit, x, xgt, valid, L):
    model.zero_grad(set_to_none=True)
    out = model(x, torch.zeros(x.shape[0], 4), True)
    loss = sum(crit(out[l], xgt[:, l], valid[:, l])[0] for l in range(L)) / L
    loss.backward()
    return float(loss.detach())


def _build(rank):
    from oracle import golden_check, inputs
    from lanedetection_end2end_b200.Loss_crit import backprojection_loss
    L, B = 2, 2
    model, args = golden_check.build_net(L, 2, 0.3, B)
    sd = model.state_dict()
    for k, v in inputs.make_erfnet_params(3, L, seed=21).items():     # every rank builds the same replica
        sd[k] = torch.from_numpy(v)
    model.load_state_dict(sd)
    model = model.cuda().train()
    for m in model.modules():
        if hasattr(m, "dropout"):
            m.dropout.p = 0
    return model, backprojection_loss(args), L, B


def _shard(rank, B):
    from oracle import inputs
    x = torch.from_numpy(inputs.make_images(B, 256, 512, seed=40 + rank)).cuda()
    xgt_np, valid_np = inputs.make_loss_targets(B, 4, seed=50 + rank)
    return x, torch.from_numpy(xgt_np).cuda(), torch.from_numpy(valid_np).cuda()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from lanedetection_end2end_b200.ddp import FlatGradAllReduce, broadcast_parameters
    model, crit, L, B = _build(rank)
    broadcast_parameters(model)
    red = FlatGradAllReduce(model)
    loss = _one_step(model, crit, *_shard(rank, B), L)
    flat = red().clone()                                   # averaged over the two ranks
    torch.cuda.synchronize()
    if rank == 0:
        q.put((loss, flat.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_mean_of_single_rank_steps():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss0, flat2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # the same two shards, one after the other, in THIS process (no process group): mean of the flat gradients
    from lanedetection_end2end_b200.ddp import FlatGradAllReduce
    model, crit, L, B = _build(0)
    flats = []
    for r in range(2):
        loss = _one_step(model, crit, *_shard(r, B), L)
        if r == 0:
            assert abs(loss - loss0) <= 1e-12 * abs(loss0)
        flats.append(FlatGradAllReduce(model)().clone().cpu().numpy())
    want = 0.5 * (flats[0] + flats[1])
    scale = np.abs(want).max()
    assert np.abs(flat2 - want).max() <= 1e-6 * scale, np.abs(flat2 - want).max() / scale
