"""Multi-GPU multichannelrx: one process per GPU over torch.distributed.

A single wideband stream is split two ways, with ONE exchange in between:

  stage 1  NCO + analysis channelizer   sharded over TIME.  Rank r owns blocks
           [r*T/G, (r+1)*T/G) of the call and also reads the P-1 = 13 blocks before them (halo:
           the filter memory) and starts its NCO at theta0 + offset*dtheta, which is exact because
           the phase is a uint32 accumulator.
  exchange all-to-all of the channelizer output: rank r keeps channels [r*N/G, (r+1)*N/G) for
           all time.  The stage-1 kernel already writes [channel][time], so the send buffer is G
           contiguous slabs and needs no repacking (NCCL over NVLink; 4 bytes per wideband sample).
  stage 2  per-channel OFDM synchroniser + packet decode, sharded over CHANNELS.  A rank feeds the G
           received slabs to its synchronisers in time order; their state carries over between
           launches, so this equals one launch over the whole time axis.
  gather   decoded frame records and payload bytes to rank 0 (variable size).

The arithmetic is done by the CUDA library through the C ABI (capi.py); this module is plumbing.
`plan()` and `exchange()` are pure tensor/index logic and are also exercised on CPU with the gloo
backend (tests/test_sharded_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist

HALO_BLOCKS = 13           # taps per branch - 1 of the receive filterbank (m = 7)


def plan(total_blocks, world):
    """time shards of a call: list of (first_block, n_blocks) per rank, equal sizes required"""
    if total_blocks % world:
        raise ValueError("the number of blocks per call (%d) must divide by the number of ranks (%d)" % (total_blocks, world))
    t = total_blocks // world
    return [(r * t, t) for r in range(world)]


def exchange(send, world, group=None):
    """send: [N, T_local] channelizer output of this rank's time shard (channel-major, contiguous).
    Returns recv: [world, N/world, T_local] = for every source rank (= time shard, in time order)
    the slab of this rank's channels."""
    N, T = send.shape
    if N % world:
        raise ValueError("the number of channels (%d) must divide by the number of ranks (%d)" % (N, world))
    recv = torch.empty((world, N // world, T), dtype=send.dtype, device=send.device)
    if world == 1:
        recv[0].copy_(send)
        return recv
    flat_s = torch.view_as_real(send).reshape(-1) if send.is_complex() else send.reshape(-1)
    flat_r = torch.view_as_real(recv).reshape(-1) if recv.is_complex() else recv.reshape(-1)
    try:
        dist.all_to_all_single(flat_r, flat_s, group=group)
    except RuntimeError:
        # backends without all_to_all (gloo on some builds): pairwise exchange
        n = flat_s.numel() // world
        rank = dist.get_rank(group)
        reqs = []
        for peer in range(world):
            if peer == rank:
                flat_r[peer * n:(peer + 1) * n].copy_(flat_s[peer * n:(peer + 1) * n])
            else:
                reqs.append(dist.isend(flat_s[peer * n:(peer + 1) * n].contiguous(), peer, group=group))
                reqs.append(dist.irecv(flat_r[peer * n:(peer + 1) * n], peer, group=group))
        for q in reqs:
            q.wait()
    return recv


def gather_frames(recs, payloads, world, rank, device, group=None):
    """variable-size gather of frame records (numpy structured array) and payload bytes to rank 0"""
    if world == 1:
        return [recs], [payloads]
    sizes = torch.tensor([recs.nbytes, len(payloads)], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    cap = int(all_sizes.sum(axis=1).max())
    buf = torch.zeros(max(cap, 1), dtype=torch.uint8, device=device)
    mine = np.concatenate([recs.view(np.uint8).reshape(-1), np.asarray(payloads, np.uint8)])
    if len(mine):
        buf[:len(mine)] = torch.from_numpy(mine).to(device)
    out = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None, None
    all_recs, all_pl = [], []
    for r in range(world):
        b = out[r].cpu().numpy()
        nr, npay = int(all_sizes[r][0]), int(all_sizes[r][1])
        all_recs.append(b[:nr].view(recs.dtype).copy())
        all_pl.append(b[nr:nr + npay].copy())
    return all_recs, all_pl


class ShardedMultichannelRx:
    """multichannelrx over `world` GPUs (one per process): time-sharded channelizer, all-to-all,
    channel-sharded synchronisers.  Every call processes `blocks_per_call` blocks of 2N wideband
    samples in total; rank r is handed its own time shard (plus halo) already in device memory."""

    def __init__(self, num_channels, M, cp_len, taper_len, blocks_per_call, rank, world, device=0):
        from . import capi
        self.N, self.K, self.M, self.cp = num_channels, 2 * num_channels, M, cp_len
        self.rank, self.world = rank, world
        if num_channels % world:
            raise ValueError("channels must divide by ranks")
        self.shards = plan(blocks_per_call, world)
        self.t_local = self.shards[rank][1]
        self.device = torch.device("cuda", device)
        # stage 1 uses the channelizer of a full N-channel handle; stage 2 a synchroniser bank of N/G streams
        self.chan = capi.MultichannelRx(num_channels, M, cp_len, taper_len, device=device, max_batch=4 * self.K)
        self.sync = capi.OfdmSync(M, cp_len, taper_len, streams=num_channels // world, device=device,
                                  max_batch=self.t_local)
        self.send = torch.empty((num_channels, self.t_local), dtype=torch.complex64, device=self.device)
        self.blocks_done = 0

    def execute_device(self, x_shard):
        """x_shard: complex64 device tensor of (HALO_BLOCKS + t_local) * 2N samples: this rank's time
        shard of the call preceded by its halo"""
        first = self.blocks_done + self.shards[self.rank][0]
        offset = (first - HALO_BLOCKS) * self.K
        self.chan.channelize_device(x_shard.data_ptr(), self.t_local, offset, self.send.data_ptr(), self.t_local)
        torch.cuda.synchronize(self.device)              # stage-1 stream -> NCCL stream
        recv = exchange(self.send, self.world)
        torch.cuda.synchronize(self.device)
        for src in range(self.world):                    # time order
            self.sync.execute_device(recv[src].data_ptr(), self.t_local, self.t_local)
        self.blocks_done += self.t_local * self.world

    def poll(self):
        recs, pl = self.sync.poll()
        recs["channel"] += self.rank * (self.N // self.world)
        return recs, pl

    def close(self):
        self.chan.close()
        self.sync.close()
