"""timing probe of the sharded receiver under torchrun: per-call time without / with the gather, P2P copy bandwidth"""
import importlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sh = importlib.import_module("liquid-usrp_b200.sharded")
w = bench.WORKLOAD
period, expected, flen = bench.make_period()
K = 2 * w["N"]
per_chunk, steps = 13, 7
tc = per_chunk * flen
rx = sh.ShardedRx(w["N"], w["M"], w["cp"], w["taper"], tc, steps, rank, world, device=local)
tile = np.concatenate([period[-sh.HALO_BLOCKS * K:], np.tile(period, per_chunk)])
d_x = torch.from_numpy(tile.view(np.float32)).cuda()
ptrs = [d_x.data_ptr()] * steps
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for it in range(6):
    sync(); t0 = time.perf_counter()
    rx.execute_device(ptrs)
    sync(); t1 = time.perf_counter()
    recs, pl = rx.poll_view()
    t = rx.gather_async(48 << 20)
    res = rx.gather_wait(t)
    sync(); t2 = time.perf_counter()
    if rank == 0: print("call %d: execute %.2f ms, gather %.2f ms, frames %d" % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), len(recs)), flush=True)
# gather pieces
import ctypes as C
capi = rx.capi
cap = 48 << 20
g = rx._g
for it in range(3):
    sync(); t0 = time.perf_counter()
    nr, nb = C.c_size_t(0), C.c_size_t(0)
    with torch.cuda.stream(rx.s2):
        capi._check(rx.L.b2_mcrx_shard_pack_results(rx.h, C.c_void_p(g["send"].data_ptr()), cap, C.byref(nr), C.byref(nb)))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    with torch.cuda.stream(rx.s2):
        g["sizes"][0] = nr.value; g["sizes"][1] = nb.value
        sz = [torch.zeros_like(g["sizes"]) for _ in range(world)]
        if world > 1: dist.all_gather(sz, g["sizes"])
        all_sizes = torch.stack(sz).cpu().numpy()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    used = int((all_sizes[:, 0] * 88 + all_sizes[:, 1]).max()); used = (used + 255) & ~255
    with torch.cuda.stream(rx.s2):
        outl = [g["recv"][0][r, :used] for r in range(world)] if rank == 0 else None
        if world > 1: dist.gather(g["send"][:used], outl, dst=0)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if rank == 0:
        g["host"][0][:, :used].copy_(g["recv"][0][:, :used], non_blocking=True)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    if rank == 0:
        for r in range(world):
            g["host"][0][r, :used].copy_(g["recv"][0][r, :used], non_blocking=True)
    torch.cuda.synchronize(); t5 = time.perf_counter()
    if rank == 0: print("pack %.2f sizes %.2f gather %.2f d2h-2d %.2f d2h-rows %.2f ms (used %d)" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), used), flush=True)
# stage 1 alone
L = rx.L
import ctypes as C
for it in range(3):
    sync(); t0 = time.perf_counter()
    with torch.cuda.stream(rx.s1):
        for i in range(steps):
            L.b2_mcrx_shard_stage1(rx.h, C.c_void_p(d_x.data_ptr()), 1000 + i)
    sync(); t1 = time.perf_counter()
    if rank == 0: print("stage 1 x%d alone: %.2f ms" % (steps, 1e3 * (t1 - t0)), flush=True)
if world > 1:
    a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % local)
    b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % ((local + 1) % world))
    for it in range(3):
        sync(); t0 = time.perf_counter()
        b.copy_(a); torch.cuda.synchronize(); t1 = time.perf_counter()
        if rank == 0: print("peer copy 256 MiB: %.1f GB/s" % (0.268 / (t1 - t0)), flush=True)
    if rank == 0: os.system("nvidia-smi topo -m | head -12")
rx.close()
if world > 1: dist.destroy_process_group()
