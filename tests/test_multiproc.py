"""CPU test (gloo, world_size 2) of the multi-GPU host logic bench.py uses: stream-per-rank sharding with no data-path
collective, barrier + MAX-over-ranks timing, whole-job aggregate computed on rank 0."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, time
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# each rank owns whole streams / closed GOPs: GOP g -> rank g % world (SURVEY 8e); no tensor crosses ranks in the data path
gops = list(range(8))
mine = [g for g in gops if g % world == rank]
frames = sum(8 for _ in mine)
dist.barrier()
t0 = time.perf_counter(); time.sleep(0.02 * (rank + 1)); ms = (time.perf_counter() - t0) * 1e3
t = torch.tensor([ms], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
tot = torch.tensor([frames], dtype=torch.int64)
dist.all_reduce(tot, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"max_ms": float(t[0]), "own_ms": ms, "frames_total": int(tot[0]), "mine": mine}))
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding_and_max_over_ranks(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29517", str(w)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["frames_total"] == 64 and d["mine"] == [0, 2, 4, 6]
    assert d["max_ms"] >= d["own_ms"] and d["max_ms"] >= 35.0       # rank 1 slept 40 ms: the MAX is what gets reported


EXCH_WORKER = r'''
import os, sys, json
sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "svt-hevc_b200"))
import torch, torch.distributed as dist
from gop_exchange import GopExchange, gops_of_rank
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
fpg, cap, rounds = 4, 4096, 3
ex = GopExchange(rank, world, torch.device("cpu"), fpg, cap)
def fill(gop, frame, nb):          # the bytes picture `frame` of GOP `gop` produces: a pattern only that (gop, frame) has
    return ((torch.arange(nb, dtype=torch.int64) * 7 + gop * 31 + frame * 5) % 251).to(torch.uint8)
ok = True
for r in range(rounds):
    gop = r * world + rank
    sizes = [100 + 37 * ((gop * fpg + f) % 11) for f in range(fpg)]          # data-dependent lengths
    for f, nb in enumerate(sizes):
        ex.frame_view(r & 1, f)[:nb] = fill(gop, f, nb)
    ex.exchange(r, sizes)
    ex.wait_all()
    if rank == 0:                   # the entropy-coding rank reads every GOP of the round, in coding order, and finds exactly what its owner wrote
        for src in range(world):
            g = r * world + src
            for f in range(fpg):
                nb = 100 + 37 * ((g * fpg + f) % 11)
                got = ex.payload_on_rank0(g, f)
                ok = ok and got.numel() == nb and bool(torch.equal(got, fill(g, f, nb)))
if rank == 0:
    print(json.dumps({"ok": ok, "order": [g for g, _, _ in ex.assembled], "owners": [s for _, s, _ in ex.assembled], "mine": gops_of_rank(0, world, rounds * world)}))
dist.barrier()
dist.destroy_process_group()
'''


def test_gop_boundary_exchange_reassembles_coding_order_on_rank0(tmp_path):
    """world-size-2 gloo run of the GOP-boundary exchange bench.py --config 4 uses over NCCL: GOP g -> rank g % world, variable-length per-picture
    payloads gathered to rank 0, which must see the GOPs in coding order with their owners' exact bytes"""
    w = tmp_path / "exch_worker.py"
    w.write_text(EXCH_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29519", str(w)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env={**os.environ, "OMP_NUM_THREADS": "1", "REPO_ROOT": ROOT})
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ok"] and d["order"] == list(range(6)) and d["owners"] == [0, 1, 0, 1, 0, 1] and d["mine"] == [0, 2, 4]
