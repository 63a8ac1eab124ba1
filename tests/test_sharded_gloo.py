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


# ---------------------------------------------------------------- the pipelined split (ShardedRx / b2_mcrx_shard_*)
def rr_worker(rank, port, ret):
    """round-robin time chunks: rank r channelizes chunks r, r + world, ... (each with its 13-block halo and the NCO phase
    of its absolute position), writes them into columns [r * tc, (r + 1) * tc) of every owner's slot -- here through a gloo
    all-to-all instead of peer stores -- and every rank then holds ITS channels over `world` consecutive chunks of time"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    x = make_stream()
    H, tc, steps = sh.HALO_BLOCKS, 24, 4                               # 4 steps x 2 ranks x 24 blocks = 2 T blocks
    pad = np.concatenate([np.zeros(H * K, np.complex64), x])
    cpp = N // WORLD
    rows = []
    for step in range(steps):
        g = sh.chunk_of(step, rank, WORLD)
        seg = pad[g * tc * K:(g * tc + H + tc) * K]
        tile = torch.from_numpy(oracle_channelize(seg, (g * tc - H) * K, tc, H))      # [N, tc]
        recv = sh.exchange(tile, WORLD)                                 # [source rank, my channels, tc]
        slot = np.zeros((cpp, WORLD * tc), np.complex64)
        for src in range(WORLD):
            slot[:, src * tc:(src + 1) * tc] = recv[src].numpy()         # the column range the stage-1 kernel writes
        rows.append(slot)
    mine = np.concatenate(rows, axis=1)
    full = oracle_channelize(pad[:(H + 2 * T) * K], -H * K, 2 * T, H)
    ret[rank] = bool(np.array_equal(mine, full[rank * cpp:(rank + 1) * cpp]))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_chunks_equal_single_process_channelizer():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=rr_worker, args=(r, port, ret)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_step_dependencies_keep_the_exchange_slots_safe():
    """replay ShardedRx's stream waits (sharded.step_dependencies) against random kernel durations on `world` ranks: a slot
    is never overwritten before every rank has read it, never read before every rank has written it, nothing deadlocks"""
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    rng = np.random.default_rng(3)
    for world, steps, slots in ((2, 7, 4), (8, 7, 4), (8, 12, 4), (4, 5, 4)):
        deps = sh.step_dependencies(steps, slots)
        for trial in range(20):
            dur = {k: rng.uniform(0.05, 1.0, (world, steps)) * rng.choice([1.0, 5.0], (world, 1)) for k in ("stage1", "stage2")}
            end = {k: np.full((world, steps), np.nan) for k in ("stage1", "stage2")}
            start = {k: np.full((world, steps), np.nan) for k in ("stage1", "stage2")}
            join = np.full((world, steps), np.nan)
            bar = np.full(steps, np.nan)                                   # completion of barrier(i) (the same on every rank)

            def ready(r, kind, i):
                ts = []
                for k2, j in deps[i][kind]:
                    v = bar[j] if k2 == "barrier" else end[k2][r, j]
                    if np.isnan(v):
                        return None
                    ts.append(v)
                return max(ts) if ts else 0.0
            progress = True
            while progress:
                progress = False
                for r in range(world):
                    for kind in ("stage1", "stage2"):
                        for i in range(steps):
                            if not np.isnan(end[kind][r, i]):
                                continue
                            prev = 0.0 if i == 0 else end[kind][r, i - 1]      # stream order
                            t = ready(r, kind, i)
                            if t is None or np.isnan(prev):
                                break
                            start[kind][r, i] = max(t, prev)
                            end[kind][r, i] = start[kind][r, i] + dur[kind][r, i]
                            progress = True
                    for i in range(steps):
                        if np.isnan(join[r, i]):
                            prev = 0.0 if i == 0 else bar[i - 1]
                            t = ready(r, "barrier", i)
                            if t is None or np.isnan(prev):
                                break
                            join[r, i] = max(t, prev)
                            progress = True
                for i in range(steps):
                    if np.isnan(bar[i]) and not np.isnan(join[:, i]).any():
                        bar[i] = join[:, i].max() + 0.02
                        progress = True
            assert not np.isnan(end["stage2"]).any(), "deadlock"
            for i in range(steps):
                assert (start["stage2"][:, i] >= end["stage1"][:, i].max() - 1e-12).all()              # read after every write
                if i >= slots:
                    assert (start["stage1"][:, i] >= end["stage2"][:, i - slots].max() - 1e-12).all()    # write after every read


def test_pack_parsing():
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    recs = np.zeros(3, pkg.FRAME_DTYPE)
    recs["channel"] = [5, 6, 7]
    recs["payload_len"] = 4
    recs["payload_offset"] = [0, 16, 32]
    pl = np.arange(48, dtype=np.uint8)
    row = np.zeros(4096, np.uint8)
    row[:32].view(np.uint64)[:] = [3, 48, 9, 0]
    row[32:32 + recs.nbytes] = recs.view(np.uint8).reshape(-1)
    row[32 + recs.nbytes:32 + recs.nbytes + 48] = pl
    r, p, tag = sh.parse_pack(row, pkg.FRAME_DTYPE)
    assert tag == 9 and list(r["channel"]) == [5, 6, 7] and np.array_equal(p, pl)
    assert sh.chunk_of(3, 1, 8) == 25
