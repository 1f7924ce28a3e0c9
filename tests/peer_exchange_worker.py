"""Worker of tests/test_gpu_peer.py (one process per GPU under torch.distributed.run): GOP hand-off to rank 0 by peer-memory copies."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "svt-hevc_b200"))
from gop_exchange import GopExchange  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fpg, cap = 4, 1 << 20
    ex = GopExchange(rank, world, dev, fpg, cap, peer=True)
    comm = torch.cuda.Stream(device=dev)
    ok = True
    for rnd in range(3):
        sizes = [((rank + 1) * 1000 + 37 * f + 11 * rnd) * 16 for f in range(fpg)]
        for f in range(fpg):      # picture f of this rank's GOP: a pattern that names rank, round and picture
            v = ex.frame_view(rnd & 1, f)
            v[:sizes[f]] = torch.arange(sizes[f], device=dev, dtype=torch.int64).mul_(rank + 3).add_(7 * f + rnd).remainder_(251).to(torch.uint8)
        e = torch.cuda.Event(); e.record(torch.cuda.current_stream()); comm.wait_event(e)
        with torch.cuda.stream(comm):
            ex.exchange(rnd, sizes)
            ex.wait_all()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            for src in range(world):
                g = rnd * world + src
                for f in range(fpg):
                    n = ((src + 1) * 1000 + 37 * f + 11 * rnd) * 16
                    want = torch.arange(n, device=dev, dtype=torch.int64).mul_(src + 3).add_(7 * f + rnd).remainder_(251).to(torch.uint8)
                    got = ex.payload_on_rank0(g, f)
                    torch.cuda.synchronize()
                    ok = ok and got.numel() == n and bool(torch.equal(got, want))
        dist.barrier()
    if rank == 0:
        order = [g for g, _, _ in ex.assembled]
        ok = ok and order == list(range(3 * world))
        print("PEER_EXCHANGE_OK" if ok else "PEER_EXCHANGE_MISMATCH", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
