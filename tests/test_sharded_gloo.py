"""Host-side logic of the multi-GPU path (liquid-usrp_b200/sharded.py) on CPU: two processes over
torch.distributed/gloo.  The CUDA kernels are replaced by the oracle's channelizer so that the
sharding arithmetic itself is pinned: time shards + 13-block halo + exact NCO phase offset, the
all-to-all layout, and the variable-size gather of frame records."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc
from b2 import pkg  # noqa: F401  (puts the package on sys.path)
import importlib

WORLD = 2
N, K, T = 4, 8, 96              # channels, filterbank size, blocks per call


def oracle_channelize(x, first_sample, n_blocks, halo_blocks):
    """x: samples starting at the halo; returns [N, n_blocks] (halo outputs dropped)"""
    L = orc.lib()
    q = L.firpfbch_crcf_create_kaiser(0, K, 7, 60.0)
    off = np.float32(-0.5 * (N - 1) / N * np.pi)
    u = L.orc_nco_constrain(off)
    n = (np.arange(len(x), dtype=np.int64) + first_sample).astype(np.uint64)
    th = ((n * np.uint64(u)) & np.uint64(0xffffffff)).astype(np.uint32).astype(np.int32)
    t = (th.astype(np.float64) * (np.pi / 2147483648.0)).astype(np.float32)
    xm = (x * (np.cos(t.astype(np.float64)) - 1j * np.sin(t.astype(np.float64)))).astype(np.complex64)
    out = np.zeros((halo_blocks + n_blocks, K), np.complex64)
    y = np.zeros(K, np.complex64)
    for b in range(halo_blocks + n_blocks):
        xi = np.ascontiguousarray(xm[b * K:(b + 1) * K])
        L.firpfbch_crcf_analyzer_execute(q, xi.ctypes.data, y.ctypes.data)
        out[b] = y
    L.firpfbch_crcf_destroy(q)
    return np.ascontiguousarray(out[halo_blocks:, :N].T)


def make_stream():
    rng = np.random.default_rng(42)
    return (rng.standard_normal(2 * T * K) + 1j * rng.standard_normal(2 * T * K)).astype(np.complex64)


def worker(rank, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    x = make_stream()
    H = sh.HALO_BLOCKS
    pad = np.concatenate([np.zeros(H * K, np.complex64), x])          # stream preceded by silence
    rows = []
    for call in range(2):                                             # two consecutive calls of T blocks
        first, tl = sh.plan(T, WORLD)[rank]
        b0 = call * T + first                                         # absolute first block of my shard
        seg = pad[b0 * K:(b0 + H + tl) * K]                           # halo + shard
        send = torch.from_numpy(oracle_channelize(seg, (b0 - H) * K, tl, H))
        recv = sh.exchange(send, WORLD)
        assert recv.shape == (WORLD, N // WORLD, tl)
        rows.append(torch.cat([recv[s] for s in range(WORLD)], dim=1).numpy())
    mine = np.concatenate(rows, axis=1)                               # my channels, all time
    full = oracle_channelize(pad[:(H + 2 * T) * K], -H * K, 2 * T, H)
    c0 = rank * (N // WORLD)
    ok_chan = np.array_equal(mine, full[c0:c0 + N // WORLD])
    # gather of variable-size frame records
    recs = np.zeros(rank + 1, pkg.FRAME_DTYPE)
    recs["channel"] = rank
    recs["payload_len"] = 3
    pl = np.arange(3 * (rank + 1), dtype=np.uint8) + 10 * rank
    all_recs, all_pl = sh.gather_frames(recs, pl, WORLD, rank, torch.device("cpu"))
    ok_gather = True
    if rank == 0:
        ok_gather = (len(all_recs) == WORLD and [len(r) for r in all_recs] == [1, 2] and
                     int(all_recs[1]["channel"][0]) == 1 and np.array_equal(all_pl[1], np.arange(6, dtype=np.uint8) + 10))
    ret[rank] = bool(ok_chan and ok_gather)
    dist.barrier()
    dist.destroy_process_group()


def test_time_shard_exchange_equals_single_process_channelizer():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=worker, args=(r, port, ret)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_plan_rejects_ragged_shards():
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    assert sh.plan(96, 4) == [(0, 24), (24, 24), (48, 24), (72, 24)]
    with pytest.raises(ValueError):
        sh.plan(97, 4)
