"""Multi-GPU multichannelrx: one process per GPU over torch.distributed.

ShardedRx (bottom of this file) is the production path: round-robin time chunks, the channelizer scattering
straight into peer GPU memory (b2_mcrx_shard_* in include/b200_ofdm.h), one one-word NCCL all-reduce per step
as the only cross-rank synchronisation, stages of successive steps overlapping on two streams.
ShardedMultichannelRx below is the first, unpipelined formulation (contiguous time shards, NCCL all-to-all); it
is kept because its pure index logic (plan / exchange / gather_frames) is what the gloo tests exercise on CPU.

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


PACK_HEADER = 32           # B2_SHARD_PACK_HEADER: uint64 n_recs, n_payload_bytes, tag, 0


def chunk_of(step, rank, world):
    """absolute index of the stream chunk rank `rank` channelizes in step `step` (chunks are dealt round-robin)"""
    return step * world + rank


def parse_pack(row, frame_dtype):
    """one rank's pack [n_recs, n_payload_bytes, tag, 0 | records | payload bytes] (uint8 array) -> (records, payloads, tag)"""
    nr, nb, tag = (int(v) for v in row[:24].view(np.uint64))
    isz = frame_dtype.itemsize
    return row[PACK_HEADER:PACK_HEADER + nr * isz].view(frame_dtype), row[PACK_HEADER + nr * isz:PACK_HEADER + nr * isz + nb], tag


def step_dependencies(steps, slots):
    """the stream-ordering rules of one call as data, for every step i: what stage 1, the barrier and stage 2 of the step
    wait for (beyond the order of their own streams).  ShardedRx.execute_* issue exactly these waits; the CPU test
    replays them against random kernel durations and checks that no exchange slot is overwritten before every rank has
    read it, and that nobody reads a slot before every rank has written it."""
    deps = []
    for i in range(steps):
        deps.append({"stage1": [("barrier", i - 2)] if i >= 2 else [],
                     "barrier": [("stage1", i)] + ([("stage2", i - (slots - 2))] if i - (slots - 2) >= 0 else []),
                     "stage2": [("barrier", i)]})
    return deps


class ShardedRx:
    """one rank of a multichannelrx spread over `world` GPUs (one process each).

    The wideband stream is cut into chunks of `chunk_blocks` blocks of 2N samples; chunk g goes to rank g % world
    (the way a front end would deal DMA buffers round), so after every STEP -- one chunk per rank -- each rank holds
    its N/world channels for `world` consecutive chunks of time and runs its synchronisers over them while the
    next step is being channelized.  Data plane: the stage-1 kernel stores every channel's run directly into the
    owning GPU's exchange buffer (CUDA IPC mapping over NVLink).  Control plane: a one-word all-reduce per step on
    the stage-1 stream (everybody's stores of the step have landed / everybody is done with the slot that comes
    round next), torch.distributed gather of the frame records at the end of a call.
    """
    SLOTS = 4
    PACK_HEADER = 32
    import os as _os
    _dbg_nocopy = _os.environ.get("B2_DBG_NOCOPY") == "1"       # probe knob: packs leave the device without their body

    def __init__(self, num_channels, M, cp_len, taper_len, chunk_blocks, steps_per_call, rank, world, device=0, group=None,
                 host_results=True):
        import ctypes as C
        from . import capi
        self.C, self.capi = C, capi
        self.L = capi.lib()
        self.N, self.K, self.rank, self.world, self.group = num_channels, 2 * num_channels, rank, world, group
        self.tc, self.steps = chunk_blocks, steps_per_call
        self.device = torch.device("cuda", device)
        self.s1 = torch.cuda.Stream(device=self.device)
        self.s2 = torch.cuda.Stream(device=self.device)
        h = C.c_void_p()
        capi._check(self.L.b2_mcrx_shard_create(num_channels, M, cp_len, taper_len, None, device, rank, world, chunk_blocks,
                                                 steps_per_call, C.c_void_p(self.s1.cuda_stream), C.c_void_p(self.s2.cuda_stream),
                                                 C.byref(h)))
        self.h = h
        if not host_results:                           # frames are collected with gather_async / gather_wait only
            capi._check(self.L.b2_mcrx_shard_host_results(self.h, 0))
        if world > 1:
            mine = (C.c_ubyte * 64)()
            capi._check(self.L.b2_mcrx_shard_export(self.h, mine))
            t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.device)
            allh = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allh, t, group=group)
            blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            capi._check(self.L.b2_mcrx_shard_connect(self.h, buf, world))
            dist.barrier(group=group)
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.step = 0                                  # steps since the stream began
        self.s3 = torch.cuda.Stream(device=self.device)          # the per-step barrier (does not hold stage 1 up)
        self.ev_stage1 = [torch.cuda.Event() for _ in range(steps_per_call)]
        self.ev_ready = [torch.cuda.Event() for _ in range(steps_per_call)]
        self.ev_done = [torch.cuda.Event() for _ in range(steps_per_call)]

    def execute_device(self, chunk_ptrs):
        """one call = steps_per_call steps; chunk_ptrs[i]: device pointer to this rank's chunk of step i, preceded
        by its 13 halo blocks ((13 + chunk_blocks) * 2N complex64 samples, 16-byte aligned)"""
        L, C, capi = self.L, self.C, self.capi
        assert len(chunk_ptrs) == self.steps
        cur = torch.cuda.current_stream(self.device)
        self.s1.wait_stream(cur)
        self.s2.wait_stream(cur)
        capi._check(L.b2_mcrx_shard_begin(self.h))
        for i in range(self.steps):
            with torch.cuda.stream(self.s1):
                if self.world > 1 and i >= 2:
                    # slot (i % SLOTS) was read by stage 2 of step i - SLOTS: barrier(i - 2), which every rank joined
                    # behind its own stage 2 of step i - SLOTS, has completed
                    self.s1.wait_event(self.ev_ready[i - 2])
                capi._check(L.b2_mcrx_shard_stage1(self.h, C.c_void_p(int(chunk_ptrs[i])), self.step))
                self.ev_stage1[i].record(self.s1)
            self._barrier(i)
            with torch.cuda.stream(self.s2):
                self.s2.wait_event(self.ev_ready[i])
                capi._check(L.b2_mcrx_shard_stage2(self.h, self.step))
                self.ev_done[i].record(self.s2)
            self.step += 1
        capi._check(L.b2_mcrx_shard_end(self.h))       # host waits for stage 2 of every step
        self._end_of_call(cur)

    def _end_of_call(self, cur):
        # the next call's first stage 1 may overwrite a slot only after the slowest rank's last stage 2
        if self.world > 1:
            with torch.cuda.stream(self.s3):
                self.s3.wait_event(self.ev_done[self.steps - 1])
                dist.all_reduce(self.flag, group=self.group)
            self.s1.wait_stream(self.s3)
        cur.wait_stream(self.s1)
        cur.wait_stream(self.s2)
        cur.wait_stream(self.s3)

    def _barrier(self, i):
        """barrier(i) on its own stream: everybody's stage 1 of step i has landed (joined behind the own stage 1), and --
        joined behind the own stage 2 of step i - SLOTS + 2 -- everybody is done with the slot that step i + 2 overwrites"""
        with torch.cuda.stream(self.s3):
            self.s3.wait_event(self.ev_stage1[i])
            if self.world > 1:
                j = i - (self.SLOTS - 2)
                if j >= 0:
                    self.s3.wait_event(self.ev_done[j])
                dist.all_reduce(self.flag, group=self.group)
            self.ev_ready[i].record(self.s3)

    def execute_host(self, host_chunks):
        """the same call fed from PINNED HOST memory: host_chunks[i] is a pinned uint8/float32/complex64 tensor holding
        [13 halo blocks | chunk] of step i; the copies go through two device buffers on a copy stream, so the H2D of
        step i + 1 overlaps stage 1 of step i"""
        if not hasattr(self, "_hbuf"):
            n = host_chunks[0].numel()
            self._hbuf = [torch.empty(n, dtype=host_chunks[0].dtype, device=self.device) for _ in range(2)]
            self._hcopy = torch.cuda.Stream(device=self.device)
            self._hev_copied = [torch.cuda.Event() for _ in range(2)]
            self._hev_used = [torch.cuda.Event() for _ in range(2)]
            self._hused = [False, False]
        L, C, capi = self.L, self.C, self.capi
        assert len(host_chunks) == self.steps
        cur = torch.cuda.current_stream(self.device)
        self.s1.wait_stream(cur); self.s2.wait_stream(cur); self._hcopy.wait_stream(cur)
        capi._check(L.b2_mcrx_shard_begin(self.h))
        for i in range(self.steps):
            b = i & 1
            with torch.cuda.stream(self._hcopy):
                if self._hused[b]:
                    self._hcopy.wait_event(self._hev_used[b])           # stage 1 that read this buffer last
                self._hbuf[b].copy_(host_chunks[i], non_blocking=True)
                self._hev_copied[b].record(self._hcopy)
            with torch.cuda.stream(self.s1):
                self.s1.wait_event(self._hev_copied[b])
                if self.world > 1 and i >= 2:
                    self.s1.wait_event(self.ev_ready[i - 2])
                capi._check(L.b2_mcrx_shard_stage1(self.h, C.c_void_p(self._hbuf[b].data_ptr()), self.step))
                self._hev_used[b].record(self.s1)
                self._hused[b] = True
                self.ev_stage1[i].record(self.s1)
            self._barrier(i)
            with torch.cuda.stream(self.s2):
                self.s2.wait_event(self.ev_ready[i])
                capi._check(L.b2_mcrx_shard_stage2(self.h, self.step))
                self.ev_done[i].record(self.s2)
            self.step += 1
        capi._check(L.b2_mcrx_shard_end(self.h))
        self._end_of_call(cur)

    def poll(self):
        C, capi = self.C, self.capi
        n, nb = C.c_size_t(), C.c_size_t()
        capi._check(self.L.b2_mcrx_shard_poll(self.h, None, 0, C.byref(n), None, 0, C.byref(nb)))
        recs = np.zeros(n.value, dtype=capi.FRAME_DTYPE)
        pl = np.zeros(max(nb.value, 1), dtype=np.uint8)
        if n.value:
            capi._check(self.L.b2_mcrx_shard_poll(self.h, C.c_void_p(recs.ctypes.data), n.value, C.byref(n),
                                                  C.c_void_p(pl.ctypes.data), len(pl), C.byref(nb)))
        return recs, pl[:nb.value]

    def poll_view(self):
        """zero-copy: arrays alias the library's pinned buffers, valid until the next call on this handle"""
        C, capi = self.C, self.capi
        pr, pp, n, nb = C.c_void_p(), C.c_void_p(), C.c_size_t(0), C.c_size_t(0)
        capi._check(self.L.b2_mcrx_shard_poll_view(self.h, C.byref(pr), C.byref(n), C.byref(pp), C.byref(nb)))
        if n.value == 0:
            return np.zeros(0, capi.FRAME_DTYPE), np.zeros(0, np.uint8)
        recs = np.frombuffer((C.c_char * (n.value * capi.FRAME_DTYPE.itemsize)).from_address(pr.value), dtype=capi.FRAME_DTYPE)
        pl = np.frombuffer((C.c_char * max(nb.value, 1)).from_address(pp.value), dtype=np.uint8)[:nb.value] if nb.value else np.zeros(0, np.uint8)
        return recs, pl

    # ---- gather of the decoded frames on rank 0: NCCL from device memory, one D2H on rank 0 that overlaps the next call
    def gather_async(self, cap_bytes, via="nccl"):
        """start collecting the frames of the call that just ended on rank 0.  Every rank packs
        [sizes | records in callback order | payloads] from device memory (cap_bytes, the same on every rank; the sizes
        travel inside the pack, so there is no size exchange and no host synchronisation), then
          via="nccl": NCCL gathers the packs on rank 0 over NVLink and rank 0 copies them to pinned host memory -- every
                      byte crosses rank 0's PCIe link;
          via="shm":  every rank copies its pack over ITS OWN PCIe link into a host buffer that rank 0 has mapped too
                      (POSIX shared memory registered with CUDA): nothing funnels through one link.
        Returns a ticket for gather_wait(); the copies overlap the next call."""
        C, capi = self.C, self.capi
        if getattr(self, "_g", None) is None or self._g["cap"] != cap_bytes or self._g["via"] != via:
            self._gather_release()
            self._g = {"cap": cap_bytes, "via": via, "send": torch.empty(cap_bytes, dtype=torch.uint8, device=self.device),
                       "copy": torch.cuda.Stream(device=self.device), "k": 0, "ev": [torch.cuda.Event() for _ in range(2)]}
            if via == "shm":
                self._shm_setup(cap_bytes)
            elif self.rank == 0:
                self._g["recv"] = [torch.empty((self.world, cap_bytes), dtype=torch.uint8, device=self.device) for _ in range(2)]
                self._g["host"] = [torch.empty((self.world, cap_bytes), dtype=torch.uint8).pin_memory() for _ in range(2)]
        g = self._g
        nr, nb = C.c_size_t(0), C.c_size_t(0)
        k = g["k"] & 1
        g["k"] += 1
        seq = g["k"]                                  # tag of this pack (1, 2, ...)
        with torch.cuda.stream(self.s2):
            if g.get("last_ev") is not None:          # the previous pack has left the send buffer
                self.s2.wait_event(g["last_ev"])
            capi._check(self.L.b2_mcrx_shard_pack_results(self.h, C.c_void_p(g["send"].data_ptr()), cap_bytes, seq, C.byref(nr), C.byref(nb)))
            src = None
            if via == "nccl":
                if self.world == 1:
                    src = g["send"].view(1, -1)
                else:
                    outl = [g["recv"][k][r] for r in range(self.world)] if self.rank == 0 else None
                    dist.gather(g["send"], outl, dst=0, group=self.group)
                    src = g["recv"][k] if self.rank == 0 else None
            done = torch.cuda.Event()
            done.record(self.s2)
        H = self.PACK_HEADER
        used = H + nr.value * capi.FRAME_DTYPE.itemsize + nb.value
        if via == "shm":
            with torch.cuda.stream(g["copy"]):
                g["copy"].wait_event(done)
                # body first, header (with the tag rank 0 polls) last: copies of one stream land in order
                cs = C.c_void_p(g["copy"].cuda_stream)
                if self._dbg_nocopy:
                    used = H
                for o in range(H, used, 4 << 20):      # in pieces: small copies of other streams slip in between
                    capi._check(self.L.b2_memcpy_async(C.c_void_p(g["mine"][k].data_ptr() + o), C.c_void_p(g["send"].data_ptr() + o),
                                                       min(4 << 20, used - o), cs))
                capi._check(self.L.b2_memcpy_async(C.c_void_p(g["mine"][k].data_ptr()), C.c_void_p(g["send"].data_ptr()), H, cs))
                g["ev"][k].record(g["copy"])
            g["last_ev"] = g["ev"][k]
            return (k, seq)
        if self.rank != 0:
            return None
        with torch.cuda.stream(g["copy"]):
            g["copy"].wait_event(done)
            for r in range(self.world):              # (row by row: a strided 2-D D2H copy is 30x slower)
                g["host"][k][r].copy_(src[r], non_blocking=True)
            g["ev"][k].record(g["copy"])
        g["last_ev"] = g["ev"][k]
        return (k, seq)

    def _gather_release(self):
        # registered host memory must be unregistered before it is unmapped (the address range may be handed out again)
        g = getattr(self, "_g", None)
        if g and g.get("mine"):
            torch.cuda.synchronize(self.device)
            for t in g["mine"]:
                torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
        self._g = {"cap": -1, "via": None}

    def _shm_setup(self, cap):
        import os
        g = self._g
        tag = "%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid() if self.world > 1 else os.getpid())
        path = lambda r, k: "/dev/shm/b2_gather_%s_r%d_%d" % (tag, r, k)
        g["mine"] = []
        for k in range(2):
            with open(path(self.rank, k), "wb") as f:
                f.truncate(cap)
            t = torch.from_file(path(self.rank, k), shared=True, size=cap, dtype=torch.uint8)
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), cap, 0)
            if int(rc) != 0:
                raise RuntimeError("cudaHostRegister failed: %s" % rc)
            g["mine"].append(t)
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.rank == 0:
            g["all"] = [[g["mine"][k] if r == 0 else torch.from_file(path(r, k), shared=True, size=cap, dtype=torch.uint8)
                         for r in range(self.world)] for k in range(2)]
        if self.world > 1:
            dist.barrier(group=self.group)
        for k in range(2):                               # the mappings keep the memory alive
            os.unlink(path(self.rank, k))

    def gather_wait(self, ticket):
        """rank 0: -> list over source ranks of (records, payload bytes) numpy views into host memory (valid until the
        gather_wait after next); other ranks: None.  With via="shm" rank 0 polls the tag every rank's copy writes last
        (no collective, nobody else waits)."""
        if ticket is None:
            return None
        k, seq = ticket
        g = self._g
        if g["via"] == "shm":
            if self.rank != 0:
                return None
            import time
            rows = [t.numpy() for t in g["all"][k]]
            t0 = time.perf_counter()
            for row in rows:
                tag = row[16:24].view(np.uint64)
                while int(tag[0]) != seq:
                    if time.perf_counter() - t0 > 30.0:
                        raise RuntimeError("timed out waiting for a rank's frames in shared memory")
                    time.sleep(0)
        else:
            g["ev"][k].synchronize()
            host = g["host"][k].numpy()
            rows = [host[r] for r in range(self.world)]
        return [parse_pack(row, self.capi.FRAME_DTYPE)[:2] for row in rows]

    def reset(self):
        self.capi._check(self.L.b2_mcrx_shard_reset(self.h))

    def close(self):
        if self.h:
            torch.cuda.synchronize(self.device)
            self._gather_release()
            self.L.b2_mcrx_shard_destroy(self.h)
            self.h = None
