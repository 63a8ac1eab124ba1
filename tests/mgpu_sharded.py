"""multi-GPU parity check of the sharded receiver, run under torchrun with 2+ ranks (one GPU each):
every rank feeds its round-robin chunks of ONE oracle-generated wideband stream; the records gathered on rank 0 must
equal those of a single-GPU multichannelrx over the same stream.  Prints "sharded ok" on rank 0."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def canonical(recs, pl):
    """records in callback order (completion index, then channel) with their payloads re-packed in that order"""
    order = np.lexsort((recs["channel"], recs["complete_index"]))
    out = recs[order].copy()
    chunks, off = [], 0
    for i in range(len(out)):
        n = int(out["payload_len"][i]) if out["header_valid"][i] else 0
        o = int(out["payload_offset"][i])
        chunks.append(pl[o:o + n])
        out["payload_offset"][i] = off
        off += n
    return out, (np.concatenate(chunks) if chunks else np.zeros(0, np.uint8))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from b2 import pkg
    import test_gpu_parity as T
    sh = importlib.import_module("liquid-usrp_b200.sharded")
    ok = True
    for name in ("c5_shape_32ch_qam64", "c2_8ch_h128"):
        case = T.CASES[name]
        N, M, cp, taper = case[:4]
        K = 2 * N
        x = T.make_input(case)
        steps, calls = 2, 2
        tc = (len(x) // K) // (steps * calls * world)
        x = x[:tc * steps * calls * world * K]
        rx = sh.ShardedRx(N, M, cp, taper, tc, steps, rank, world, device=local)
        bufs = T.sharded_chunks(x, K, tc, steps * calls, rank, world, sh.HALO_BLOCKS)
        got_r, got_p = [], []
        for c in range(calls):
            rx.execute_device([b.data_ptr() for b in bufs[c * steps:(c + 1) * steps]])
            rx.poll_view()
            # both routes to rank 0's host memory: NCCL gather + rank 0's D2H, and every rank's own D2H into shared memory
            res = rx.gather_wait(rx.gather_async(1 << 24, via=("nccl" if c == 0 else "shm")))
            if rank == 0:
                for r_, p_ in res:
                    r_ = r_.copy()
                    r_["payload_offset"] += sum(len(q) for q in got_p)
                    got_r.append(r_); got_p.append(p_.copy())
        rx.close()
        if rank == 0:
            g = pkg.MultichannelRx(N, M, cp, taper, device=local)
            g.execute(x)
            fo, po = g.poll()
            g.close()
            fg, pg = canonical(np.concatenate(got_r), np.concatenate(got_p))
            fo, po = canonical(fo, po)
            try:
                T.assert_frames_equal(fo, po, fg, pg)
                print("%s: %d frames equal" % (name, len(fo)), flush=True)
            except AssertionError as e:
                ok = False
                print("MISMATCH %s: %s" % (name, e), flush=True)
    dist.barrier()
    if rank == 0 and ok:
        print("sharded ok", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
