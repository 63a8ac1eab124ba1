"""timing probe of the sharded receiver under torchrun: per-call time of execute alone, with the gather (nccl / shm)
overlapped as in bench.py, and of stage 1 alone.  B2_SHARD_COPY=1 selects the copy-engine exchange."""
import ctypes as C, importlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    bench.bind_near_gpu(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sh = importlib.import_module("liquid-usrp_b200.sharded")
w = bench.WORKLOAD
period, expected, flen = bench.make_period()
K = 2 * w["N"]
per_chunk, steps = 13, int(os.environ.get("STEPS", "7"))
tc = per_chunk * flen
rx = sh.ShardedRx(w["N"], w["M"], w["cp"], w["taper"], tc, steps, rank, world, device=local, host_results=False)
tile = np.concatenate([period[-sh.HALO_BLOCKS * K:], np.tile(period, per_chunk)])
d_x = torch.from_numpy(tile.view(np.float32)).cuda()
ptrs = [d_x.data_ptr()] * steps
cap = int(1.03 * w["N"] * steps * per_chunk * (88 + w["payload"] + 16)) + (1 << 20)
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def say(*a):
    if rank == 0: print(*a, flush=True)
for it in range(3): rx.execute_device(ptrs)
sync(); t0 = time.perf_counter()
for it in range(6): rx.execute_device(ptrs)
sync(); say("execute alone: %.2f ms per call (%s exchange)" % (1e3 * (time.perf_counter() - t0) / 6, "copy-engine" if os.environ.get("B2_SHARD_COPY") == "1" else "fused peer-store"))
for via in ("shm", "nccl"):
    pend = None
    for it in range(3):
        rx.execute_device(ptrs); t = rx.gather_async(cap, via=via)
        if pend is not None: rx.gather_wait(pend)
        pend = t
    sync(); t0 = time.perf_counter()
    for it in range(6):
        rx.execute_device(ptrs); t = rx.gather_async(cap, via=via)
        if pend is not None: res = rx.gather_wait(pend)
        pend = t
    res = rx.gather_wait(pend)
    sync(); dt = (time.perf_counter() - t0) / 6
    say("execute + gather via %s: %.2f ms per call -> %.1f GS/s, frames on rank 0: %d" % (via, 1e3 * dt, steps * tc * K * world / dt / 1e9, sum(len(r) for r, _ in res) if res else -1))
L = rx.L
for it in range(2):
    sync(); t0 = time.perf_counter()
    with torch.cuda.stream(rx.s1):
        for i in range(steps): L.b2_mcrx_shard_stage1(rx.h, C.c_void_p(d_x.data_ptr()), 1000 + i)
    sync(); say("stage 1 x%d alone: %.2f ms" % (steps, 1e3 * (time.perf_counter() - t0)))
rx.close()
if world > 1: dist.destroy_process_group()
