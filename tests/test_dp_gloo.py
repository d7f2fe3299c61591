"""CPU, world_size 2 over gloo: the N>1 host logic of the path — stream
partitioning, bucketed asynchronous gradient all-reduce launched from post-accumulate-grad
hooks, end-of-backward join, global-'mean' loss equivalence, no_sync accumulation.
The wrapped module here is the CPU oracle network (tests may use oracle/); the CUDA module
itself cannot run without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from oracle import lucy_oracle as LO
from statecatcher_b200.dp import StreamDataParallel, partition_streams, shard_loss_scale


def test_partition_streams():
    assert [list(partition_streams(512, 8, r))[:1] + [len(partition_streams(512, 8, r))] for r in (0, 7)] == [[0, 64], [448, 64]]
    parts = [partition_streams(10, 4, r) for r in range(4)]
    assert sorted(i for p in parts for i in p) == list(range(10))
    assert [len(p) for p in parts] == [3, 3, 2, 2]
    with pytest.raises(ValueError):
        partition_streams(4, 2, 2)


class _OracleNet(nn.Module):
    """LucyRNN + CTC on the CPU via the oracle's closed form, parameters as nn.Parameters
    (including the dead W_r gate rows that never receive a gradient)."""

    def __init__(self, cfg, seed):
        super().__init__()
        self.cfg = cfg
        P = LO.random_params(cfg, seed, dtype=torch.float64)
        self.names = list(P)
        self.params = nn.ParameterList([nn.Parameter(P[k]) for k in self.names])

    def forward(self, x, state=None):
        P = dict(zip(self.names, self.params))
        return LO.forward_closed(P, self.cfg, x, state)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, overlap=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        cfg = LO.OracleConfig(input_dim=6, hidden_dim=8, num_layers=2, vocab_size=7, fused_ops=False, layer_norm=True)
        net = _OracleNet(cfg, 5)
        if overlap:
            ddp = StreamDataParallel(net, bucket_mb=0.0005, overlap=True)   # tiny buckets -> several all-reduces, launched from hooks
            assert len(ddp.buckets) > 3
        else:
            ddp = StreamDataParallel(net)                          # default: one flat buffer, exchanged after the backward
            assert len(ddp.buckets) == 1 and not ddp.overlap
        g = torch.Generator().manual_seed(11)
        B, T = 4, 9                                                # global batch: 4 streams, 2 per rank
        xs = [torch.randn(B, T, 6, generator=g, dtype=torch.float64) for _ in range(2)]
        toks = [torch.randint(1, 7, (B, 3), generator=g) for _ in range(2)]
        inl, tgl = [T, T - 2, T, T], [3, 2, 3, 1]
        mine = partition_streams(B, world, rank)
        sl = slice(mine.start, mine.stop)
        crit = nn.CTCLoss(blank=0, zero_infinity=True)
        state = None
        for i in range(2):                                         # two carried segments
            if state:
                state = LO.detach_states(state)
            logits, state = ddp(xs[i][sl], state)
            loss = crit(logits.log_softmax(-1).transpose(0, 1), toks[i][sl], inl[sl], tgl[sl])
            loss.backward()                                        # grads accumulate over segments
        assert ddp.n_allreduce >= 3 if overlap else ddp.n_allreduce == 1
        grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in zip(net.names, net.params)}
        # no_sync: local accumulation only
        for p in net.params:
            p.grad = None
        with ddp.no_sync():
            logits, _ = ddp(xs[0][sl])
            crit(logits.log_softmax(-1).transpose(0, 1), toks[0][sl], inl[sl], tgl[sl]).backward()
        local_only = net.params[0].grad.clone()
        if rank == 0:
            torch.save({"grads": grads, "state_h": [t.detach() for t in state[0]], "local_only": local_only}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True], ids=["after-backward", "hook-overlapped"])
def test_stream_data_parallel_matches_single_process(tmp_path, overlap):
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, overlap), nprocs=2, join=True)
    got = torch.load(out)
    # single process over the concatenated batch (reduction='mean' over the GLOBAL batch)
    cfg = LO.OracleConfig(input_dim=6, hidden_dim=8, num_layers=2, vocab_size=7, fused_ops=False, layer_norm=True)
    net = _OracleNet(cfg, 5)
    g = torch.Generator().manual_seed(11)
    B, T = 4, 9
    xs = [torch.randn(B, T, 6, generator=g, dtype=torch.float64) for _ in range(2)]
    toks = [torch.randint(1, 7, (B, 3), generator=g) for _ in range(2)]
    inl, tgl = [T, T - 2, T, T], [3, 2, 3, 1]
    crit = nn.CTCLoss(blank=0, zero_infinity=True)
    state = None
    for i in range(2):
        if state:
            state = LO.detach_states(state)
        logits, state = net(xs[i], state)
        crit(logits.log_softmax(-1).transpose(0, 1), toks[i], inl, tgl).backward()
    for k, p in zip(net.names, net.params):
        if p.grad is None:
            assert got["grads"][k] is None, k                       # dead r-gate params stay grad-less
            continue
        torch.testing.assert_close(got["grads"][k], p.grad, rtol=1e-9, atol=1e-12, msg=k)
    # rank 0 carried only ITS streams' state: equals rows 0..1 of the single-process state
    for a, b in zip(got["state_h"], state[0]):
        torch.testing.assert_close(a, b[:2].detach(), rtol=1e-9, atol=1e-12)
    # no_sync grad differs from the synchronised one (it is the local shard's only)
    assert not torch.allclose(got["local_only"], got["grads"][net.names[0]])


# ---- unequal shards: 5 streams over 2 ranks (3 + 2), local losses weighted by shard_loss_scale ----
def _uneven_problem():
    cfg = LO.OracleConfig(input_dim=6, hidden_dim=8, num_layers=1, vocab_size=7, fused_ops=True, layer_norm=False)
    g = torch.Generator().manual_seed(23)
    B, T = 5, 8
    x = torch.randn(B, T, 6, generator=g, dtype=torch.float64)
    toks = torch.randint(1, 7, (B, 3), generator=g)
    return cfg, B, T, x, toks, [T, T, T - 3, T, T - 1], [3, 1, 2, 3, 2]


def _uneven_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        cfg, B, T, x, toks, inl, tgl = _uneven_problem()
        net = _OracleNet(cfg, 9)
        ddp = StreamDataParallel(net)
        mine = partition_streams(B, world, rank)
        sl = slice(mine.start, mine.stop)
        logits, _ = ddp(x[sl])
        loss = nn.CTCLoss(blank=0, zero_infinity=True)(logits.log_softmax(-1).transpose(0, 1), toks[sl], inl[sl], tgl[sl])
        (loss * shard_loss_scale(len(mine), B, world)).backward()
        if rank == 1:
            torch.save({k: (p.grad.clone() if p.grad is not None else None) for k, p in zip(net.names, net.params)}, out)
    finally:
        dist.destroy_process_group()


def test_unequal_shards_reproduce_the_global_mean(tmp_path):
    assert shard_loss_scale(64, 512, 8) == 1.0 and shard_loss_scale(3, 5, 2) == 1.2 and shard_loss_scale(0, 5, 2) == 0.0
    with pytest.raises(ValueError):
        shard_loss_scale(6, 5, 2)
    out = str(tmp_path / "rank1.pt")
    mp.spawn(_uneven_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    cfg, B, T, x, toks, inl, tgl = _uneven_problem()
    net = _OracleNet(cfg, 9)
    logits, _ = net(x)
    nn.CTCLoss(blank=0, zero_infinity=True)(logits.log_softmax(-1).transpose(0, 1), toks, inl, tgl).backward()
    for k, p in zip(net.names, net.params):
        if p.grad is None:
            assert got[k] is None, k
            continue
        torch.testing.assert_close(got[k], p.grad, rtol=1e-9, atol=1e-12, msg=k)
