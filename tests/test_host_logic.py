"""CPU: host-side logic of the mirror — batch rule, sharding, the world_size-2 gather (gloo), scheduler host state."""
from __future__ import annotations

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fourierdiffusion_b200 as fd
from fourierdiffusion_b200 import distributed as fdd
from fourierdiffusion_b200.engine import model_kind_of, renorm_fixed_point, scheduler_params
from oracle import fdiff_oracle as O


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 1000, 4096):
        for ws in (1, 2, 3, 4, 8):
            spans = [fdd.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == fdd.shard_sizes(n, ws)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank: int, ws: int, port: int, n_total: int):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        full = torch.arange(n_total * 6, dtype=torch.float32).reshape(n_total, 3, 2)
        lo, hi = fdd.shard_range(n_total, rank, ws)
        got = fdd.all_gather_series(full[lo:hi].clone(), n_total)
        assert torch.equal(got, full), f"rank {rank}"
        assert fdd.world() == (rank, ws)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])  # equal shards -> one all_gather_into_tensor; ragged -> padded gather
def test_all_gather_series_world_size_2_gloo(n_total):
    mp.spawn(_gather_worker, args=(2, _free_port(), n_total), nprocs=2, join=True)


def test_scheduler_host_state_matches_oracle():
    for fourier in (False, True):
        for L in (7, 24, 256):
            s = fd.VPScheduler(fourier_noise_scaling=fourier)
            s.set_noise_scaling(L)
            assert torch.equal(s.G, O.g_vector(L, fourier))
            s.set_timesteps(1000)
            ts, dt = O.make_timesteps(1000)
            assert torch.equal(s.timesteps, ts) and torch.equal(s.step_size, dt)
    assert scheduler_params(fd.VPScheduler(beta_min=0.1, beta_max=20.0)) == (0, 0.1, 20.0)
    assert scheduler_params(fd.VEScheduler(sigma_min=0.01, sigma_max=2.0)) == (1, 0.01, 2.0)
    with pytest.raises(NotImplementedError):
        scheduler_params(object())


def test_model_kind_and_state_dict_keys():
    sch = fd.VPScheduler()
    t = fd.ScoreModule(n_channels=3, max_len=20, noise_scheduler=sch, d_model=8, n_head=4, num_layers=2)
    l = fd.LSTMScoreModule(n_channels=3, max_len=20, noise_scheduler=sch, d_model=8, num_layers=2)
    m = fd.MLPScoreModule(n_channels=3, max_len=20, noise_scheduler=sch, d_model=8, d_mlp=16, num_layers=2)
    assert [model_kind_of(x) for x in (t, l, m)] == [0, 1, 2]
    assert "backbone.layers.1.self_attn.in_proj_weight" in t.state_dict() and "pos_encoder.embedding.weight" in t.state_dict()
    assert "backbone.1.weight_hh_l0" in l.state_dict() and "pos_encoder.embedding.weight" not in l.state_dict()
    assert "backbone.1.3.weight" in m.state_dict() and m.state_dict()["embedder.weight"].shape == (8, 60)
    with pytest.raises(NotImplementedError):
        model_kind_of(torch.nn.Linear(2, 2))


def test_renorm_fixed_point_equals_oracle_and_is_idempotent():
    torch.manual_seed(1)
    E = torch.randn(64, 16) * 1.5
    a = renorm_fixed_point(E, 4.0)
    assert float(a.norm(dim=1).max()) <= 4.0 + 1e-5
    assert torch.allclose(a, O.renorm_positional_table(E, 4.0), atol=1e-6)
    assert torch.equal(renorm_fixed_point(a, 4.0), a)


def test_sampler_rejects_unknown_scheduler():
    class Fake:
        noise_scheduler = object()
        n_channels, max_len = 1, 4

    with pytest.raises(NotImplementedError, match="Scheduler not recognized"):
        fd.DiffusionSampler(score_model=Fake(), sample_batch_size=2)


def _stack_table(B, L, layers, lag):
    import ctypes as C

    from fourierdiffusion_b200 import _lib

    lib = _lib.load()
    n = lib.fd_stack_task_table(B, L, layers, lag, None, 0)
    assert n > 0
    buf = (C.c_uint32 * n)()
    assert lib.fd_stack_task_table(B, L, layers, lag, C.cast(buf, C.c_void_p), n) == n
    return list(buf)


@pytest.mark.parametrize("B,L", [(1, 256), (2, 256), (3, 252), (7, 187), (16, 32), (5, 50), (64, 256), (33, 100), (9, 129)])
@pytest.mark.parametrize("lag", [-2, -1, 0, 1, 5, 10**6])
def test_stack_task_queue_is_a_topological_order(B, L, lag):
    """The persistent encoder-stack kernel (csrc/fd_step.cu) claims tasks in queue order and a task only waits for tasks that are
    already claimed — so the queue must list every dependency before its dependant, and every task exactly once."""
    layers = 3
    M = B * L
    n_tiles = (M + 127) // 128
    table = _stack_table(B, L, layers, B // 2 if lag == -1 else lag)
    assert len(table) == layers * (4 * B + n_tiles)
    att_done, ffn_done = {}, {}
    seen = set()
    for e in table:
        assert e not in seen
        seen.add(e)
        is_ffn, layer, idx = e >> 31, (e >> 24) & 0x7F, e & 0xFFFFFF
        if is_ffn:
            m = idx
            assert m < n_tiles and layer < layers
            s_first, s_last = (m * 128) // L, min(m * 128 + 127, M - 1) // L
            for b in range(s_first, s_last + 1):  # needs all 4 head groups of every series the tile touches, this layer
                assert att_done.get((layer, b), 0) == 4, (e, b)
            for b in range(s_first, s_last + 1):
                ffn_done[(layer, b)] = ffn_done.get((layer, b), 0) + 1
        else:
            b, g = idx >> 2, idx & 3
            assert b < B and layer < layers
            t_first, t_last = (b * L) // 128, ((b + 1) * L - 1) // 128
            if layer > 0:  # needs every FFN tile that covers the series at the previous layer
                assert ffn_done.get((layer - 1, b), 0) == t_last - t_first + 1, (e, b)
            att_done[(layer, b)] = att_done.get((layer, b), 0) + 1
