"""Closed-GOP sharding over ranks and the GOP-boundary exchange (closed GOPs: Source/API/EbApi.h:203-220, intraRefreshType >= 0).

No picture of GOP k references GOP k +- 1 and ME is open loop, so GOP g is encoded whole by rank g % world.  The ONLY data-path communication is at
the GOP boundary: the per-picture outputs the entropy coder consumes (packed quantised coefficients + TU offsets + nz counts + ME / OIS / SAO records)
go from the rank that produced them to the entropy-coding rank (0), device to device (NCCL over NVLink on GPUs; gloo on CPU in the tests), while the
sender already runs its next GOP.  Rank 0 reassembles the GOPs in coding order.

One round = every rank has finished one GOP.  Sizes differ per picture (the coefficient stream is data dependent), so a round starts with an
all_gather of the byte counts, then point-to-point transfers of exactly those bytes.
"""
import torch
import torch.distributed as dist


def gop_owner(gop_index, world):
    return gop_index % world


def gops_of_rank(rank, world, n_gops):
    return [g for g in range(n_gops) if gop_owner(g, world) == rank]


class GopExchange:
    """frames_per_gop pictures per GOP; every rank owns a send buffer laid out [frame][frame_capacity bytes] per parity (double buffered: GOP k + 1 is
    produced while GOP k is in flight); rank 0 owns one receive buffer per (source rank, parity)."""

    def __init__(self, rank, world, device, frames_per_gop, frame_capacity, comm_stream=None, peer=False):
        """peer=True (GPUs of one node): rank 0's receive buffers are mapped into every producer's address space (CUDA IPC through the C ABI's
        svtb200_peer_* entry points) and a GOP's pictures move by device-to-device copies over NVLink on the copy engines; torch.distributed then only
        carries the byte counts.  peer=False: point-to-point send / recv of the same bytes (NCCL on GPUs, gloo in the CPU tests)."""
        self.rank, self.world, self.device = rank, world, device
        self.fpg, self.cap = frames_per_gop, int(frame_capacity)
        self.send = [torch.zeros(self.fpg * self.cap, dtype=torch.uint8, device=device) for _ in range(2)]
        self.recv = {}
        self.peer = bool(peer) and world > 1
        self.remote = [None, None]          # producers: rank 0's buffer for this rank, per parity (device pointers valid in this process)
        self.recv_ptr = {}                  # rank 0: (source rank, parity) -> device pointer of the exported buffer
        if self.peer:
            import svtb200 as S
            self.S = S
            handles = [None]
            if rank == 0:
                table = {}
                for src in range(1, world):
                    for parity in range(2):
                        ptr = S.peer_alloc(self.fpg * self.cap)
                        self.recv_ptr[(src, parity)] = ptr
                        table[(src, parity)] = S.peer_export(ptr)
                handles = [table]
            dist.broadcast_object_list(handles, src=0)
            if rank != 0:
                self.remote = [S.peer_open(handles[0][(rank, parity)]) for parity in range(2)]
        elif rank == 0:
            for src in range(1, world):
                self.recv[src] = [torch.zeros(self.fpg * self.cap, dtype=torch.uint8, device=device) for _ in range(2)]
        self.comm_stream = comm_stream
        self.assembled = []          # rank 0: (gop_index, source rank, [bytes per frame]) in coding order
        self.pending = []
        self.bytes_moved = 0

    def frame_view(self, parity, frame):
        """the bytes picture `frame` of the GOP under construction writes its outputs into"""
        return self.send[parity][frame * self.cap:(frame + 1) * self.cap]

    def exchange(self, round_index, frame_bytes):
        """frame_bytes: this rank's used bytes per frame of the GOP it just finished (list of fpg ints).  Posts the transfers of the round (asynchronous;
        wait_all() completes them) and, on rank 0, records the coding order.  Returns the bytes this rank sends (or, on rank 0, receives)."""
        parity = round_index & 1
        mine = torch.tensor(frame_bytes, dtype=torch.int64, device=self.device)
        sizes = torch.zeros(self.world * self.fpg, dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_gather_into_tensor(sizes, mine)
        else:
            sizes.copy_(mine)
        sizes = sizes.cpu().view(self.world, self.fpg).tolist()
        ops, moved = [], 0
        if self.peer:
            # the data plane: this rank's pictures straight into rank 0's memory, on the caller's current stream (stream-ordered after the kernels that
            # produced them); rank 0 only learns the sizes
            if self.rank == 0:
                moved = sum(sum(sizes[src]) for src in range(1, self.world))
            else:
                st = torch.cuda.current_stream().cuda_stream
                for f, nb in enumerate(sizes[self.rank]):
                    if nb:
                        self.S.peer_copy_async(self.remote[parity] + f * self.cap, self.send[parity].data_ptr() + f * self.cap, nb, st)
                        moved += nb
        elif self.world > 1:
            if self.rank == 0:
                for src in range(1, self.world):
                    for f, nb in enumerate(sizes[src]):
                        if nb:
                            ops.append(dist.P2POp(dist.irecv, self.recv[src][parity][f * self.cap:f * self.cap + nb], src))
                            moved += nb
            else:
                for f, nb in enumerate(sizes[self.rank]):
                    if nb:
                        ops.append(dist.P2POp(dist.isend, self.send[parity][f * self.cap:f * self.cap + nb], 0))
                        moved += nb
        if ops:
            self.pending.extend(dist.batch_isend_irecv(ops))
        if self.rank == 0:
            for src in range(self.world):          # GOPs of this round in coding order: round * world + src
                self.assembled.append((round_index * self.world + src, src, sizes[src]))
        self.bytes_moved += moved
        return moved

    def wait_all(self):
        for w in self.pending:
            w.wait()
        self.pending = []

    def payload_on_rank0(self, gop_index, frame):
        """rank 0: the received (or locally produced) bytes of one picture -- what the entropy coder of that GOP would read"""
        for g, src, sizes in self.assembled:
            if g == gop_index:
                parity = (gop_index // self.world) & 1
                if src != 0 and self.peer:      # the exported buffer is raw device memory: hand out a copy
                    out = torch.empty(sizes[frame], dtype=torch.uint8, device=self.device)
                    self.S.peer_copy_async(out.data_ptr(), self.recv_ptr[(src, parity)] + frame * self.cap, sizes[frame], torch.cuda.current_stream().cuda_stream)
                    return out
                buf = self.send[parity] if src == 0 else self.recv[src][parity]
                return buf[frame * self.cap:frame * self.cap + sizes[frame]]
        raise KeyError(gop_index)
