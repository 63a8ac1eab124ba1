"""host-side copy bandwidth of a box with every GPU busy at once (torchrun): D2H into cudaMallocHost memory and into
registered POSIX shared memory, H2D from pinned memory; per-rank GB/s, min / max over ranks"""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
ncpu = bench.bind_near_gpu(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
path = "/dev/shm/b2_d2h_probe_%d" % rank
with open(path, "wb") as f: f.truncate(n)
shm = torch.from_file(path, shared=True, size=n, dtype=torch.uint8)
assert int(torch.cuda.cudart().cudaHostRegister(shm.data_ptr(), n, 0)) == 0
def run(dst, src, label):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    bw = torch.tensor([4 * n / dt / 1e9], device="cuda")
    lo, hi, tot = bw.clone(), bw.clone(), bw.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(tot)
    if rank == 0: print("%-28s per rank %.1f .. %.1f GB/s, box total %.1f GB/s" % (label, float(lo), float(hi), float(tot)), flush=True)
run(pinned, d, "D2H -> cudaMallocHost")
run(shm, d, "D2H -> registered /dev/shm")
run(d, pinned, "H2D <- cudaMallocHost")
if rank == 0: print("cpus bound per rank: %d, affinity now %s" % (ncpu, sorted(os.sched_getaffinity(0))[:4]), flush=True)
torch.cuda.cudart().cudaHostUnregister(shm.data_ptr())
os.unlink(path)
if world > 1: dist.destroy_process_group()
