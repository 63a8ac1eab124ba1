#!/usr/bin/env python
"""bench.py -- throughput of the multichannelrx hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of multichannelrx::Execute over one batch of synthetic wideband input
(NCO mix-down + 2N-channel polyphase analysis + N x ofdmflexframesync + packet decode + frame
records back on the host).  Workload at every N: the north-star shape, 256 channels x 512
subcarriers, cp 64, 64-QAM, no FEC, 1200-byte payloads, every channel transmitting back to back
(BASELINE.json configs[4] per-GPU share; SURVEY.md 8d "C5").  The input is ONE steady-state frame
period produced by the CPU oracle's transmitter (the reference's lib/multichanneltx.cc over
oracle/), tiled on the device; the tiling is seamless because every frame ends in > 500 samples of
silence and the NCO phase is periodic in the period length.

  value     wideband Msamples/s, input resident in HBM, whole job (all ranks)
  e2e       same through the C ABI with HOST (pinned) input: H2D + kernels + D2H of the records
  roofline  dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM peak
  cpu_baseline   the reference's lib/multichannelrx.cc over the oracle, on this box's host CPU

Multi-GPU (--gpus N under torchrun), weak scaling in both modes:
  --mode replicas (default)  every rank runs a complete 256-channel receiver on its own wideband
                             stream (the natural partition of the path: independent receivers, no
                             data-path collective; rank 0 only reduces the timings)
  --mode sharded             ONE wideband stream, N x longer: stage 1 time-sharded, NCCL all-to-all
                             of the channelizer output, stage 2 channel-sharded (256/N channels per
                             rank over the whole time axis), frames gathered to rank 0
                             (liquid-usrp_b200/sharded.py).  Stage 2 is a serial
                             recurrence per channel, so this mode trades throughput for a single
                             coherent stream; DESIGN.md discusses it.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = dict(name="multichannelrx N=256 M=512 cp=64 taper=16 qam64 fec=none payload=1200B, all channels back-to-back",
                N=256, M=512, cp=64, taper=16, payload=1200, mod="qam64", bps=6, fec0="none", nd=356)
# BASELINE.json configs[2], receive side: 64 channels, M = 256, 16-QAM, conv r1/2 K=7 (measured as an extra leg, "config64")
WORKLOAD_64 = dict(name="multichannelrx N=64 M=256 cp=32 taper=8 qam16 fec0=conv-v27 payload=1200B, all channels back-to-back",
                   N=64, M=256, cp=32, taper=8, payload=1200, mod="qam16", bps=4, fec0="v27", nd=178)
B_ALG_PATH = 16.10          # 8 B in + 4 B channelizer out + 4 B sync in + 0.10 B payload, per wideband sample
B_ALG = {"analyzer_kernel": 12.0, "sync_kernel": 4.10, "packet_decode_kernel": 0.20}
# DRAM traffic per wideband sample measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum of one --set full capture
# per kernel): read from the summary tools/ncu_traffic.py writes next to the captures it was computed from, so that it
# cannot go stale silently -- no file, no claim (traffic: null)
def traffic_per_sample():
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return {k: float(v["dram_bytes_per_sample"]) for k, v in json.load(f)["kernels"].items()}
    except Exception:
        return {}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_period(w=None):
    """one steady-state frame period of the workload from the CPU oracle transmitter
    -> (x[period] complex64, expected payload per channel, frame length in channel samples)"""
    import refmc
    w = w or WORKLOAD
    L = refmc.ref_lib()
    N, M, cp = w["N"], w["M"], w["cp"]
    enc = w["payload"] + 4
    if w["fec0"] == "v27":
        enc = (2 * (8 * enc + 6) + 7) // 8
    nd = w["nd"]
    nsym = 3 + -(-288 // nd) + -(-(-(-8 * enc // w["bps"])) // nd) + 1
    flen = nsym * (M + cp)
    mod = {"qam64": refmc.MOD_QAM64, "qam16": refmc.MOD_QAM16}[w["mod"]]
    fec0 = {"none": refmc.FEC_NONE, "v27": refmc.FEC_CONV_V27}[w["fec0"]]
    tx = refmc.McTx(L, N, M, cp, w["taper"])
    x = tx.run(2 * flen, w["payload"], mod, fec0, refmc.FEC_NONE, seed=0xB2000000, gain=1.0 / N)
    tx.close()
    K = 2 * N
    period = x[flen * K:2 * flen * K].copy()
    expected = [L.frame_data(0xB2000000, c, 1, w["payload"]) for c in range(N)]
    return period, expected, flen


class Clocks(threading.Thread):
    """sample SM clocks / throttle reasons while the timed region runs (NVML, ~2 ms per sample;
    nvidia-smi as a fallback)"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
            try:
                r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]   # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
            return [sm, self.max_sm] + flags
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        v = [s.strip() for s in out.split(",")]
        return [float(v[0]), float(v[1])] + [s.lower().startswith("active") for s in v[2:6]]

    def run(self):
        while not self.stop_flag:
            try:
                self.rows.append(self.sample())
            except Exception:
                pass
            time.sleep(0.002 if self.nvml is not None else 0.1)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_receivers(period, threads, reps, steps=None, seconds=None, warmup=1):
    """`threads` independent copies of the reference receiver (lib/multichannelrx.cc compiled unmodified
    over the oracle's restatement of liquid-dsp; the reference's DSP path is single-threaded, so a box is
    filled by independent receivers), each fed `reps` frame periods per step -> (samples, seconds, frames)"""
    import refmc
    w = WORKLOAD
    L = refmc.ref_lib()
    x = np.tile(period, reps)
    rxs = [refmc.McRx(L, w["N"], w["M"], w["cp"], w["taper"]) for _ in range(threads)]
    count = [0] * threads
    frames = [0] * threads

    def work(i, nsteps, until):
        k = 0
        while (nsteps is None or k < nsteps) and (until is None or time.perf_counter() < until):
            rxs[i].execute(x)
            fr, _pl = rxs[i].frames()
            frames[i] += len(fr)
            count[i] += len(x)
            k += 1

    def run_all(nsteps, until):
        ts = [threading.Thread(target=work, args=(i, nsteps, until)) for i in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    run_all(max(warmup, 1), None)
    count[:] = [0] * threads
    frames[:] = [0] * threads
    t0 = time.perf_counter()
    run_all(steps, None if seconds is None else t0 + seconds)
    dt = time.perf_counter() - t0
    for r in rxs:
        r.close()
    return sum(count), dt, sum(frames), len(x)


def run_reference(args, rank):
    """the reference's own CPU implementation of the path: lib/multichannelrx.cc (unmodified) over the
    oracle's restatement of liquid-dsp, one independent receiver per host core"""
    if rank != 0:
        return
    w = WORKLOAD
    period, expected, flen = make_period()
    cores = cpu_cores()
    reps = 2                                         # bounded sample: 2 frame periods per receiver per step
    n, dt, nfr, per_step = cpu_receivers(period, cores, reps, steps=args.steps, warmup=max(args.warmup, 1))
    val = n / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (oracle transmitter, one frame period tiled)",
            "config": {"workload": w["name"], "samples_per_step": per_step * cores, "frames_decoded": nfr,
                       "parallelism": "%d independent receivers, one per host core" % cores},
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port",
                             "sample": "%d wideband samples per receiver per step x %d receivers: reference lib/multichannelrx.cc "
                                       "compiled unmodified over the oracle's C restatement of liquid-dsp (liquid-dsp itself is not "
                                       "installable here), gcc -O2; the reference's DSP path has no threads, so the box is filled "
                                       "with independent receivers" % (per_step, cores)},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measure_config64(args, local_rank, pkg, barrier):
    """BASELINE.json configs[2] on the receive side (64 channels, M = 256, 16-QAM, conv r1/2 K=7), device-resident
    input, same protocol as the headline (extra leg, reported under "config64")"""
    import torch
    w = WORKLOAD_64
    period, expected, flen = make_period(w)
    reps = 48                                        # 60 M wideband samples, 481 MB per step
    n_step = len(period) * reps
    d_x = torch.from_numpy(period.view(np.float32)).cuda().repeat(reps).contiguous()
    rx = pkg.MultichannelRx(w["N"], w["M"], w["cp"], w["taper"], device=local_rank, max_batch=n_step)
    for _ in range(3):
        rx.execute_device(d_x.data_ptr(), n_step)
        recs, pl = rx.poll()
    assert len(recs) >= w["N"] * (reps - 1) and int(recs["payload_valid"].min()) == 1
    for i in (0, len(recs) // 2, len(recs) - 1):
        c, o = int(recs["channel"][i]), int(recs["payload_offset"][i])
        assert np.array_equal(pl[o:o + w["payload"]], expected[c][1])
    steps = max(3, min(args.steps, 10))
    kt = np.zeros(4)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        rx.execute_device(d_x.data_ptr(), n_step)
        kt += np.array(rx.last_timing())
        rx.poll_view()
    barrier()
    dt = time.perf_counter() - t0
    rx.close()
    kt /= steps
    return {"workload": w["name"], "value": n_step * steps / dt / 1e6, "unit": "Msamples/s", "ms_per_step": 1e3 * dt / steps,
            "samples_per_step": n_step, "steps": steps,
            "kernels_ms_per_step": {"analyzer_kernel": kt[0], "sync_kernel": kt[1], "packet_decode_kernel": kt[2], "call": kt[3]}}


def measure_tx(args, local_rank, pkg, barrier, cpu=True):
    """the transmit mirror (multichanneltx, lib/multichanneltx.cc:165-242) at the headline shape: every channel re-armed
    at each frame boundary through b2_mctx_update_many, samples left in device memory.  Extra leg, reported under "tx"."""
    import torch
    import refmc
    w = WORKLOAD
    N, M, cp, taper = w["N"], w["M"], w["cp"], w["taper"]
    K, W = 2 * N, M + cp
    nsym = 3 + -(-288 // w["nd"]) + -(-(-(-8 * (w["payload"] + 4) // w["bps"])) // w["nd"]) + 1
    calls = nsym * W                                   # GenerateSamples calls (blocks of 2N samples) per frame period
    tx = pkg.MultichannelTx(N, M, cp, taper, device=local_rank)
    rng = np.random.default_rng(7)
    payloads = rng.integers(0, 256, (N, w["payload"]), dtype=np.uint8)
    headers = rng.integers(0, 256, (N, 8), dtype=np.uint8)
    chans = np.arange(N, dtype=np.uint32)
    out = torch.empty(calls * K * 2, dtype=torch.float32, device="cuda")
    frames = max(8, args.steps)
    kt = np.zeros(4)
    for f in range(3 + frames):
        if f == 3:
            barrier()
            t0 = time.perf_counter()
            kt[:] = 0
        took = tx.update_many(chans, headers, payloads, refmc.MOD_QAM64, refmc.FEC_NONE, refmc.FEC_NONE)
        assert took == N, took
        tx.generate_device(out.data_ptr(), calls)
        kt += np.array(tx.last_timing())
    barrier()
    dt = time.perf_counter() - t0
    tx.close()
    # what was generated is a frame per channel: the receiver under test decodes them
    rx = pkg.MultichannelRx(N, M, cp, taper, device=local_rank, max_batch=calls * K)
    x = out.view(-1, 2)
    rx.execute_device((x * (1.0 / N)).contiguous().data_ptr(), calls * K)
    rx.execute_device(torch.zeros(4 * W * K * 2, dtype=torch.float32, device="cuda").data_ptr(), 4 * W * K)
    recs, pl = rx.poll()
    rx.close()
    assert len(recs) == N and int(recs["payload_valid"].min()) == 1, (len(recs),)
    c = int(recs["channel"][5]); o = int(recs["payload_offset"][5])
    assert np.array_equal(pl[o:o + w["payload"]], payloads[c])
    n = frames * calls * K
    kern_ms = float(kt[:3].sum()) / frames
    peak, _src = peaks()
    res = {"workload": "multichanneltx N=%d M=%d cp=%d qam64 payload=%dB, every channel re-armed at each frame boundary, samples left on the device" % (N, M, cp, w["payload"]),
           "value": n / dt / 1e6, "unit": "Msamples/s", "ms_per_frame_period": 1e3 * dt / frames, "samples_per_frame_period": calls * K,
           "kernels_ms_per_frame_period": kern_ms,
           "roofline": {"bound": "hbm", "alg_bytes_per_sample": 16.0, "achieved": calls * K * 16.0 / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None,
                        "peak": peak, "unit": "GB/s", "frac": (calls * K * 16.0 / (kern_ms * 1e-3) / 1e9 / peak) if kern_ms > 0 else None,
                        "note": "kernels only (packet encode + frame generator + synthesis bank / NCO); SURVEY.md 8d: 16 B per wideband output sample"}}
    if cpu:
        # the reference's lib/multichanneltx.cc over the oracle, one transmitter per host core, ~5 s
        cores = cpu_cores()
        L = refmc.ref_lib()
        counts = [0] * cores

        def work(i, until):
            t = refmc.McTx(L, N, M, cp, taper)
            while time.perf_counter() < until:
                t.run(W * 4, w["payload"], refmc.MOD_QAM64, refmc.FEC_NONE, refmc.FEC_NONE, seed=0xB2000000 + i, gain=1.0 / N)
                counts[i] += W * 4 * K
            t.close()
        tc0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=(i, tc0 + 5.0)) for i in range(cores)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        res["cpu_baseline"] = {"value": sum(counts) / (time.perf_counter() - tc0) / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                               "sample": "~5 s of the reference's lib/multichanneltx.cc (unmodified) over the oracle, one transmitter per host core"}
    return res


METRIC = "complex Msamples/s through multichannelrx (64ch OFDM) at 1/2/4/8 GPU vs CPU"


def bind_near_gpu(index):
    """run this rank's host threads (and therefore its pinned allocations, first touch) on the CPUs next to its GPU:
    without it every rank of a box pins on NUMA node 0 and the host side of the PCIe copies is shared"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_sharded(args, rank, local_rank, world, period, expected, flen):
    """ONE wideband stream over `world` GPUs (liquid-usrp_b200/sharded.py ShardedRx over b2_mcrx_shard_*): chunks of the
    stream dealt round-robin to the ranks, channelizer storing straight into the owning GPU's memory over NVLink,
    synchronisers sharded over channels, a one-word NCCL all-reduce per step, frame records gathered on rank 0.
    Weak scaling: every rank channelizes `reps` frame periods per step-of-the-bench (the N = 1 workload), so the stream
    is `world` times longer."""
    import importlib
    import torch
    import torch.distributed as dist
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    w = WORKLOAD
    K = 2 * w["N"]
    # a chunk = 13 frame periods (74 880 blocks, 38 M wideband samples), 7 chunks per rank per call = 91 periods
    per_chunk = 13
    steps = max(1, args.reps // per_chunk)
    tc = per_chunk * flen
    rx = sh.ShardedRx(w["N"], w["M"], w["cp"], w["taper"], tc, steps, rank, world, device=local_rank, host_results=False)
    # the stream is periodic in `period` and every chunk starts on a period boundary: [halo | chunk] is the same for all
    tile = np.concatenate([period[-sh.HALO_BLOCKS * K:], np.tile(period, per_chunk)])
    d_x = torch.from_numpy(tile.view(np.float32)).cuda()
    h_x = torch.from_numpy(tile.view(np.float32)).pin_memory()
    ptrs = [d_x.data_ptr()] * steps
    dev = torch.device("cuda", local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frames_call = w["N"] * steps * per_chunk                      # frames a rank decodes per call (its channels, all chunks of the call)
    cap = int(1.03 * frames_call * (88 + w["payload"] + 16)) + (1 << 20)
    pending = [None]
    gathered = [0]

    def one_call(host):
        if host:
            rx.execute_host([h_x] * steps)
        else:
            rx.execute_device(ptrs)
        recs, pl = rx.poll_view()                                 # (empty: the frames are taken from device memory below)
        ticket = rx.gather_async(cap, via=args.gather)                             # device-side ordering, NCCL gather to rank 0, D2H there: overlaps the next call
        if pending[0] is not None:
            res = rx.gather_wait(pending[0])
            if res is not None:
                gathered[0] = sum(len(r) for r, _p in res)
        pending[0] = ticket
        return recs, pl, gathered[0]

    def drain():
        if pending[0] is not None:
            res = rx.gather_wait(pending[0])
            if res is not None:
                gathered[0] = sum(len(r) for r, _p in res)
                # rank 0 holds every rank's frames: spot-check one payload per source rank
                for r, p_ in res:
                    if len(r):
                        c, o = int(r["channel"][-1]), int(r["payload_offset"][-1])
                        assert int(r["payload_valid"].min()) == 1 and int(r["header_valid"].min()) == 1
                        key = (r["complete_index"].astype(np.uint64) << np.uint64(16)) | r["channel"].astype(np.uint64)
                        assert bool(np.all(key[1:] >= key[:-1])), "records of a rank are not in callback order"
                        assert np.array_equal(p_[o:o + w["payload"]], expected[c][1]), "gathered payload mismatch on channel %d" % c
            pending[0] = None
        return gathered[0]

    n_call = steps * tc * K * world                  # wideband samples of the whole stream per call
    for _ in range(args.warmup):
        recs, pl, _n = one_call(False)
    n_got = drain()
    if rank == 0:
        assert n_got >= w["N"] * (steps * per_chunk * world - 1), "frames missing: %d" % n_got
    clk = Clocks(local_rank)
    clk.start()
    barrier()
    t0 = time.perf_counter()
    nfr = 0
    for _ in range(args.steps):
        recs, pl, n = one_call(False)
        nfr += n
    nfr += drain() - 0
    barrier()
    dt = time.perf_counter() - t0
    for _ in range(2):
        one_call(True)
    drain()
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        recs, pl, n = one_call(True)
        d2h += cap
    drain()
    barrier()
    dt_e2e = time.perf_counter() - t1
    clk.stop_flag = True
    clk.join()
    times = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    rx.close()
    peak, peak_src = peaks()
    val = n_call * args.steps / float(times[0]) / 1e6
    return {"value": val, "e2e": n_call * args.steps / float(times[1]) / 1e6, "ms_per_step": 1e3 * float(times[0]) / args.steps,
            "samples_per_step": n_call, "frames_per_step": nfr // max(args.steps, 1), "steps_per_call": steps, "chunk_blocks": tc,
            "h2d_bytes_per_step": n_call * 8 + steps * world * sh.HALO_BLOCKS * K * 8, "d2h_bytes_per_step": d2h // max(args.steps, 1) * world,
            "launches_per_step": steps * world * (1 + 4), "clocks": clk.summary(),
            "path_frac": val * 1e6 * B_ALG_PATH / 1e9 / (peak * world)}


def run_single(args, rank, local_rank, world, period, expected, flen, extras=True):
    """one complete 256-channel receiver per rank (N = 1: THE receiver; N > 1: independent replicas, no data-path
    collective) -> the JSON line as a dict"""
    import torch
    import torch.distributed as dist
    from b2 import pkg
    w = WORKLOAD
    n_step = len(period) * args.reps
    # device-resident input (the "stubbed UHD source"), tiled on the device
    d_period = torch.from_numpy(period.view(np.float32)).cuda()
    d_x = d_period.repeat(args.reps).contiguous()
    h_x = torch.from_numpy(np.tile(period, args.reps).view(np.float32)).pin_memory()
    rx = pkg.MultichannelRx(w["N"], w["M"], w["cp"], w["taper"], device=local_rank, max_batch=n_step)
    L = pkg.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def check(recs, pl, nrep):
        assert len(recs) >= w["N"] * (nrep - 1), "frames missing: %d" % len(recs)
        assert int(recs["header_valid"].min()) == 1 and int(recs["payload_valid"].min()) == 1, "invalid frame"
        for i in (0, len(recs) // 2, len(recs) - 1):
            c = int(recs["channel"][i])
            o = int(recs["payload_offset"][i])
            assert np.array_equal(pl[o:o + w["payload"]], expected[c][1]), "payload mismatch on channel %d" % c

    # ---- device-resident leg (value)
    for _ in range(args.warmup):
        rx.execute_device(d_x.data_ptr(), n_step)
        recs, pl = rx.poll()
    check(recs, pl, args.reps)
    clk = Clocks(local_rank)
    clk.start()
    kt = np.zeros(4)
    launches = 0
    barrier()
    t0 = time.perf_counter()
    nfr = 0
    t_exec = t_poll = 0.0
    n_timed = 0
    for i in range(args.steps):
        ta = time.perf_counter()
        rx.execute_device(d_x.data_ptr(), n_step)
        tb = time.perf_counter()
        if i % 2 == 0:                               # per-kernel CUDA-event sums: every other step (43 event queries each)
            kt += np.array(rx.last_timing())
            n_timed += 1
        launches += rx.last_launches()[0]
        recs, pl = rx.poll_view()
        t_poll += time.perf_counter() - tb
        t_exec += tb - ta
        nfr += len(recs)
    barrier()
    dt = time.perf_counter() - t0
    # ---- host leg (e2e): pinned host input -> C ABI -> records on the host
    for _ in range(2):
        L.b2_mcrx_execute(rx.h, C.c_void_p(h_x.data_ptr()), n_step)
        recs, pl = rx.poll()
    barrier()
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        rc = L.b2_mcrx_execute(rx.h, C.c_void_p(h_x.data_ptr()), n_step)
        assert rc == 0
        recs, pl = rx.poll_view()                    # records + decoded payloads, already DMA'd into pinned host memory
        d2h += recs.nbytes + len(pl) + 24 * rx.last_launches()[1]
    barrier()
    dt_e2e = time.perf_counter() - t1
    clk.stop_flag = True
    clk.join()
    check(recs, pl, args.reps)

    # max over ranks of the device-timed region
    times = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dt, dt_e2e = float(times[0]), float(times[1])
    total = n_step * args.steps * world
    kt_avg = kt / max(n_timed, 1)                 # ms per step summed over the chunks: analyzer, sync, decode, whole call
    names = ["analyzer_kernel", "sync_kernel", "packet_decode_kernel"]
    dom = int(np.argmax(kt_avg[:3]))
    peak, peak_src = peaks()
    nchunks = max(rx.last_launches()[1], 1)

    def roof(i):
        ach = n_step * B_ALG[names[i]] / (kt_avg[i] * 1e-3) / 1e9
        return {"alg_bytes_per_sample": B_ALG[names[i]], "launches_per_step": nchunks, "avg_launch_ms": kt_avg[i] / nchunks,
                "achieved": ach, "frac": ach / peak, "traffic": (traffic[names[i]] * n_step / nchunks) if names[i] in traffic else None}
    traffic = traffic_per_sample()
    per_kernel = {names[i]: roof(i) for i in range(3)}
    achieved = per_kernel[names[dom]]["achieved"]
    line = {"metric": METRIC, "value": total / dt / 1e6, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (oracle transmitter, one frame period tiled on device)",
            "config": {"workload": w["name"], "samples_per_step": n_step, "input_bytes_per_step": n_step * 8,
                       "l2_policy": "input (%.0f MB/step) larger than L2" % (n_step * 8 / 1e6),
                       "frames_per_step": nfr // args.steps, "parallelism": "independent receivers x%d" % world,
                       "pipeline_chunks_per_step": nchunks},
            "e2e": {"value": total / dt_e2e / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": n_step * 8,
                    "d2h_bytes_per_step": d2h // args.steps},
            "gpu_launches": launches,
            "kernels_ms_per_step": {"analyzer_kernel": kt_avg[0], "sync_kernel": kt_avg[1], "packet_decode_kernel": kt_avg[2], "call": kt_avg[3],
                                    "note": "stages of successive chunks overlap on separate streams / SM partitions; the three sums can exceed the call"},
            "host_ms_per_step": {"execute_call": 1e3 * t_exec / args.steps, "poll_call": 1e3 * t_poll / args.steps},
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": per_kernel[names[dom]]["traffic"], "peak_source": peak_src,
                         "alg_bytes_per_sample": B_ALG[names[dom]],
                         "note": "kernels of successive chunks share the SMs, so the per-kernel times are inflated by the overlap; 'path' is the whole call (DESIGN.md 6)",
                         "kernels": per_kernel,
                         "path": {"alg_bytes_per_sample": B_ALG_PATH, "achieved": n_step * B_ALG_PATH / (kt_avg[3] * 1e-3) / 1e9,
                                  "frac": n_step * B_ALG_PATH / (kt_avg[3] * 1e-3) / 1e9 / peak}},
            "clocks": clk.summary()}
    if extras and args.seconds > 0:
        clk2 = Clocks(local_rank)
        clk2.start()
        barrier()
        ts = time.perf_counter()
        k = 0
        while time.perf_counter() - ts < args.seconds:
            rx.execute_device(d_x.data_ptr(), n_step)
            rx.poll_view()
            k += 1
        barrier()
        dts = time.perf_counter() - ts
        clk2.stop_flag = True
        clk2.join()
        line["sustained"] = {"seconds": dts, "steps": k, "value": n_step * k / dts / 1e6, "unit": "Msamples/s", "clocks": clk2.summary(),
                             "note": "the same device-resident step back to back for >= --seconds"}
    if extras and not args.no_config64 and world == 1:
        line["config64"] = measure_config64(args, local_rank, pkg, barrier)
    if extras and not args.no_tx and world == 1:
        line["tx"] = measure_tx(args, local_rank, pkg, barrier, cpu=not args.no_cpu)
    if args.receivers > 1:
        rxs = [rx] + [pkg.MultichannelRx(w["N"], w["M"], w["cp"], w["taper"], device=local_rank, max_batch=n_step) for _ in range(args.receivers - 1)]

        def drive(r, k):
            for _ in range(k):
                r.execute_device(d_x.data_ptr(), n_step)
                r.poll_view()
        for phase_steps in (2, args.steps):
            ths = [threading.Thread(target=drive, args=(r, phase_steps)) for r in rxs]
            barrier()
            tm = time.perf_counter()
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            barrier()
            tm = time.perf_counter() - tm
        line["multi_receiver"] = {"receivers": args.receivers, "value": args.receivers * n_step * args.steps / tm / 1e6, "unit": "Msamples/s",
                                  "note": "aggregate of independent 256-channel receivers on ONE GPU, each fed the same device-resident stream"}
        for r in rxs[1:]:
            r.close()
    if extras and rank == 0 and world == 1 and not args.no_cpu:       # reported baseline: rank 0 at N = 1 only
        cores = cpu_cores()
        ncpu, tcpu, _nf, per_step = cpu_receivers(period, cores, 2, seconds=10.0)
        n1, t1, _nf1, _ = cpu_receivers(period, 1, 2, seconds=4.0)
        line["cpu_baseline"] = {"value": ncpu / tcpu / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                                "single_thread_value": n1 / t1 / 1e6,
                                "sample": "%d wideband samples in ~10 s over %d independent receivers (one per host core; the "
                                          "reference's DSP path is single-threaded): reference lib/multichannelrx.cc compiled "
                                          "unmodified over the oracle's C restatement of liquid-dsp, gcc -O2" % (ncpu, cores)}
    rx.close()
    return line




def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reps", type=int, default=91, help="frame periods per step (91 -> 2^28 wideband samples, the per-GPU share "
                    "of BASELINE configs[4]'s 2^31; 2.1 GB of input per step, far larger than L2)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--mode", default="auto", choices=["auto", "replicas", "sharded"], help="N > 1: auto / sharded = ONE stream split over the "
                    "GPUs (the headline; independent replicas are reported alongside), replicas = independent receivers only")
    ap.add_argument("--gather", default="shm", choices=["shm", "nccl"], help="N > 1: how the decoded frames reach rank 0's host memory: shm = every rank "
                    "over its own PCIe link into shared host memory, nccl = NCCL gather over NVLink to rank 0's GPU, then rank 0's PCIe link")
    ap.add_argument("--no-tx", action="store_true", help="skip the extra transmit (multichanneltx) leg")
    ap.add_argument("--seconds", type=float, default=2.0, help="extra leg: run the device-resident step back to back for this long and report the sustained rate and clocks under 'sustained'")
    ap.add_argument("--no-config64", action="store_true", help="skip the extra 64-channel (BASELINE configs[2]) leg")
    ap.add_argument("--receivers", type=int, default=1, help="extra leg: R independent receivers sharing this GPU (reported under "
                    "'multi_receiver', never as the headline): shows that one receiver is bound by its 256 serial chains")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from b2 import pkg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_near_gpu(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = WORKLOAD
    period, expected, flen = make_period()
    if (world > 1 and args.mode != "replicas") or args.mode == "sharded":
        sh = run_sharded(args, rank, local_rank, world, period, expected, flen)
        steps_keep = args.steps
        args.steps = max(3, min(args.steps, 6))
        rep = run_single(args, rank, local_rank, world, period, expected, flen, extras=False)
        args.steps = steps_keep
        if rank == 0:
            line = dict(rep)
            line.update({"value": sh["value"], "ms_per_step": sh["ms_per_step"], "steps": args.steps,
                         "e2e": {"value": sh["e2e"], "unit": "Msamples/s", "h2d_bytes_per_step": sh["h2d_bytes_per_step"],
                                 "d2h_bytes_per_step": sh["d2h_bytes_per_step"]},
                         "gpu_launches": sh["launches_per_step"] * args.steps, "clocks": sh["clocks"]})
            line["config"] = {"workload": WORKLOAD["name"], "samples_per_step": sh["samples_per_step"],
                              "input_bytes_per_step": sh["samples_per_step"] * 8, "l2_policy": "input per rank (%.0f MB/step) larger than L2" % (sh["samples_per_step"] * 8 / 1e6 / world),
                              "frames_per_step": sh["frames_per_step"],
                              "parallelism": "one stream split over %d GPUs: round-robin time chunks -> channelizer storing into the owning GPU over NVLink -> "
                                             "%d channels/GPU synchronisers; NCCL: one-word all-reduce per step + gather of frame records" % (world, WORKLOAD["N"] // world),
                              "steps_per_call": sh["steps_per_call"], "chunk_blocks": sh["chunk_blocks"]}
            line["roofline"] = {"bound": "hbm", "kernel": "whole path", "achieved": sh["path_frac"] * peaks()[0], "peak": peaks()[0], "unit": "GB/s",
                                "frac": sh["path_frac"], "traffic": None, "peak_source": peaks()[1], "alg_bytes_per_sample": B_ALG_PATH,
                                "note": "per-GPU average over the split: %.2f B of algorithmic traffic per wideband sample" % B_ALG_PATH}
            line["replicas"] = {"value": rep["value"], "e2e": rep["e2e"]["value"], "unit": "Msamples/s", "ms_per_step": rep["ms_per_step"],
                                "parallelism": rep["config"]["parallelism"], "note": "independent receivers, no data-path collective (round-1 headline mode)"}
            for k in ("kernels_ms_per_step", "host_ms_per_step"):
                line.pop(k, None)
            print(json.dumps(line), flush=True)
    else:
        line = run_single(args, rank, local_rank, world, period, expected, flen)
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
