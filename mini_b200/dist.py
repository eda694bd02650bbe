"""Multi-GPU BFS host layer: one process (rank) per GPU, cyclic 1D vertex partition, NCCL exchange.

The reference is single-GPU (README.md:4); this implements SURVEY.md 8e:
  push level : local LB advance that buckets remote discoveries by owner inside the kernel
               (b200_mg_bfs_push) -> alltoall(counts) -> alltoallv(vertex ids) -> label-if-unvisited
               on the receiver (b200_mg_bfs_absorb);
  pull level : allgather of the frontier-bitmap slices -> early-exit pull over the local rows
               (b200_mg_bfs_pull);
  every level: one allreduce of {|F_next|, arcs, deg(F_next)} so all ranks take the same
               push / pull / stop decision.

`DistBFS` is backend- and communicator-agnostic so the same control flow runs
  * on GPUs over NCCL (GpuRank + TorchComm)                          -- bench.py, -m gpu tests with --gpus N
  * on one GPU with P virtual ranks in threads (GpuRank + ThreadComm) -- -m gpu tests on a single GPU
  * on CPU over gloo with a NumPy stand-in for the kernels (tests/)   -- world_size-2 host-logic tests.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence

from . import partition as PT

PUSH, PULL = 0, 1


# --------------------------------------------------------------------------- communicators
class TorchComm:
    """torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, device):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device

    def all_to_all_counts(self, counts: Sequence[int]) -> List[int]:
        import torch
        send = torch.tensor(list(counts), dtype=torch.int64, device=self.device)
        recv = torch.empty_like(send)
        if self.dist.get_backend() == "gloo":
            # gloo has no all_to_all_single on every build: gather everything, pick our column
            rows = [torch.empty_like(send) for _ in range(self.world)]
            self.dist.all_gather(rows, send)
            return [int(rows[p][self.rank]) for p in range(self.world)]
        self.dist.all_to_all_single(recv, send)
        return recv.tolist()

    def all_to_all_v(self, send: List, recv: List):
        """send[p] -> rank p; recv[p] <- rank p (tensors of the exchanged counts; own entry ignored)."""
        if self.dist.get_backend() == "gloo":
            reqs = []
            for p in range(self.world):
                if p == self.rank:
                    continue
                if send[p].numel():
                    reqs.append(self.dist.isend(send[p], p))
                if recv[p].numel():
                    reqs.append(self.dist.irecv(recv[p], p))
            for r in reqs:
                r.wait()
            return
        self.dist.all_to_all(recv, send)

    def all_reduce_sum(self, vals: Sequence[int]) -> List[int]:
        import torch
        t = torch.tensor(list(vals), dtype=torch.int64, device=self.device)
        self.dist.all_reduce(t)
        return t.tolist()

    def all_gather_into(self, full, part):
        if self.dist.get_backend() == "gloo":
            chunks = list(full.chunk(self.world))
            self.dist.all_gather(chunks, part)
            return
        self.dist.all_gather_into_tensor(full, part)

    def barrier(self):
        self.dist.barrier()


class ThreadComm:
    """P virtual ranks as threads of one process (single-GPU tests of the partitioned kernels)."""

    class Shared:
        def __init__(self, world: int):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared: "ThreadComm.Shared", rank: int):
        self.s, self.rank, self.world = shared, rank, shared.world

    def _exchange(self, obj):
        self.s.slots[self.rank] = obj
        self.s.barrier.wait()
        out = list(self.s.slots)
        self.s.barrier.wait()
        return out

    def all_to_all_counts(self, counts):
        rows = self._exchange(list(counts))
        return [rows[p][self.rank] for p in range(self.world)]

    def all_to_all_v(self, send, recv):
        import torch
        torch.cuda.synchronize()
        rows = self._exchange(send)
        for p in range(self.world):
            if p != self.rank and recv[p].numel():
                recv[p].copy_(rows[p][self.rank])
        torch.cuda.synchronize()
        self.s.barrier.wait()

    def all_reduce_sum(self, vals):
        rows = self._exchange(list(vals))
        return [sum(r[i] for r in rows) for i in range(len(vals))]

    def all_gather_into(self, full, part):
        import torch
        torch.cuda.synchronize()
        rows = self._exchange(part)
        n = part.numel()
        for p in range(self.world):
            full[p * n:(p + 1) * n].copy_(rows[p])
        torch.cuda.synchronize()
        self.s.barrier.wait()

    def barrier(self):
        self.s.barrier.wait()


# --------------------------------------------------------------------------- GPU rank backend
class CMgState(C.Structure):
    _fields_ = [("rank", C.c_int32), ("num_ranks", C.c_int32), ("n_global", C.c_int64), ("n_local", C.c_int64),
                ("labels", C.c_void_p), ("known", C.c_void_p), ("frontier_bitmap", C.c_void_p),
                ("next_slice", C.c_void_p), ("box_counts", C.c_void_p)]


class GpuRank:
    """Per-rank device state + the b200_mg_* C-ABI steps.  Buffers are torch tensors so the
    collectives can use them directly."""

    def __init__(self, ctx, rank: int, world: int, n_global: int, graph_local):
        import torch
        from . import lib as L
        self.ctx, self.rank, self.world = ctx, rank, world
        self.n_global, self.n_local = n_global, n_global // world
        self.g = graph_local
        self._L = L.load_library()
        self._check = L._check
        dev = ctx.torch_device
        i32 = dict(dtype=torch.int32, device=dev)
        self.labels = torch.empty(self.n_local, **i32)
        self.known = torch.empty(n_global // 32, **i32)
        self.frontier_bitmap = torch.empty(n_global // 32, **i32)
        self.next_slice = torch.empty(self.n_local // 32, **i32)
        self.box_counts = torch.zeros(8, dtype=torch.int64, device=dev)
        self.frontier = [torch.empty(self.n_local, **i32), torch.empty(self.n_local, **i32)]
        self.sel = 0
        # remote discoveries: a rank sends each remote vertex at most once over the whole traversal
        self.boxes = [None if p == rank else torch.empty(self.n_local, **i32) for p in range(world)]
        self.inbox = torch.empty(max(self.n_local * max(world - 1, 1), 1), **i32)
        self.st = CMgState(rank, world, n_global, self.n_local, self.labels.data_ptr(), self.known.data_ptr(),
                           self.frontier_bitmap.data_ptr(), self.next_slice.data_ptr(), self.box_counts.data_ptr())
        self.cg = graph_local.cview()
        self._box_ptrs = (C.c_void_p * world)(*[None if b is None else b.data_ptr() for b in self.boxes])
        self.launches = 0

    # -- steps -----------------------------------------------------------------------------
    def init(self, src: int) -> int:
        flen = C.c_int64()
        self._check(self._L.b200_mg_bfs_init(self.ctx._h, C.byref(self.st), src, self.frontier[0].data_ptr(), C.byref(flen)),
                    "b200_mg_bfs_init")
        self.sel = 0
        return flen.value

    def push(self, level: int, flen: int):
        counts = (C.c_int64 * self.world)()
        arcs, deg = C.c_int64(), C.c_int64()
        self._check(self._L.b200_mg_bfs_push(self.ctx._h, C.byref(self.cg), C.byref(self.st), level,
                                             self.frontier[self.sel].data_ptr(), flen,
                                             self.frontier[self.sel ^ 1].data_ptr(), self._box_ptrs, self.n_local,
                                             counts, C.byref(arcs), C.byref(deg)), "b200_mg_bfs_push")
        return list(counts), arcs.value, deg.value

    def send_views(self, counts):
        e = self.inbox[:0]
        return [e if p == self.rank else self.boxes[p][:counts[p]] for p in range(self.world)]

    def recv_views(self, counts):
        out, off = [], 0
        for p in range(self.world):
            k = 0 if p == self.rank else counts[p]
            out.append(self.inbox[off:off + k])
            off += k
        return out, off

    def absorb(self, level: int, total: int):
        nlen, deg = C.c_int64(), C.c_int64()
        self._check(self._L.b200_mg_bfs_absorb(self.ctx._h, C.byref(self.cg), C.byref(self.st), level,
                                               self.inbox.data_ptr(), total, self.frontier[self.sel ^ 1].data_ptr(),
                                               C.byref(nlen), C.byref(deg)), "b200_mg_bfs_absorb")
        return nlen.value, deg.value

    def swap(self):
        self.sel ^= 1

    def pull(self, level: int):
        found, arcs, deg = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._L.b200_mg_bfs_pull(self.ctx._h, C.byref(self.cg), C.byref(self.st), level, C.byref(found),
                                             C.byref(arcs), C.byref(deg)), "b200_mg_bfs_pull")
        return found.value, arcs.value, deg.value

    def or_known(self):
        self._check(self._L.b200_mg_bitmap_or(self.ctx._h, self.known.data_ptr(), self.frontier_bitmap.data_ptr(),
                                              self.n_global // 32), "b200_mg_bitmap_or")

    def list_to_slice(self, flen: int):
        self._check(self._L.b200_mg_list_to_slice(self.ctx._h, C.byref(self.st), self.frontier[self.sel].data_ptr(), flen,
                                                  self.next_slice.data_ptr()), "b200_mg_list_to_slice")

    def slice_to_list(self) -> int:
        k = C.c_int64()
        self._check(self._L.b200_mg_slice_to_list(self.ctx._h, C.byref(self.st), self.next_slice.data_ptr(),
                                                  self.frontier[self.sel].data_ptr(), C.byref(k)), "b200_mg_slice_to_list")
        return k.value

    def gather_buffers(self):
        return self.frontier_bitmap, self.next_slice

    def reached_degree_sum(self) -> int:
        import torch
        off = self.g.row_offsets.to(torch.int64) & 0xFFFFFFFF
        deg = off[1:] - off[:-1]
        return int(deg[self.labels >= 0].sum().item())

    def reached_count(self) -> int:
        return int((self.labels >= 0).sum().item())


def build_rank_graph(ctx, scale: int, edge_factor: int, seed: int, rank: int, world: int):
    """This rank's rows of the symmetrised RMAT graph (cyclic partition), built on its own GPU."""
    import torch
    from . import lib as L
    lib = L.load_library()
    m_local = C.c_int64()
    L._check(lib.b200_rmat_part_count(ctx._h, scale, edge_factor, seed, rank, world, C.byref(m_local)), "b200_rmat_part_count")
    n_local = (1 << scale) // world
    off = torch.empty(n_local + 1, dtype=torch.int32, device=ctx.torch_device)
    idx = torch.empty(max(m_local.value, 1), dtype=torch.int32, device=ctx.torch_device)
    L._check(lib.b200_rmat_build_csr_part(ctx._h, scale, edge_factor, seed, rank, world, m_local.value, off.data_ptr(),
                                          idx.data_ptr()), "b200_rmat_build_csr_part")
    return L.Graph(n_local, m_local.value, off, idx[:m_local.value])


# --------------------------------------------------------------------------- the level loop
class DistBFS:
    """Direction-optimising BFS over P ranks.  `rank` is a GpuRank-like backend, `comm` a communicator."""

    def __init__(self, rank, comm, n_global: int, m_global: int, mode: str = "beamer", alpha: float = 15.0,
                 beta: float = 18.0):
        self.r, self.c = rank, comm
        self.n, self.m = n_global, m_global
        self.mode, self.alpha, self.beta = mode, alpha, beta
        self.levels = []

    def run(self, src: int) -> int:
        r, c = self.r, self.c
        self.levels = []
        flen_local = r.init(src)
        flen = 1
        direction, level = PUSH, 0
        m_unexplored = self.m
        while True:
            if direction == PUSH:
                counts, arcs_l, deg_l = r.push(level, flen_local)
                recv_counts = c.all_to_all_counts(counts)
                recv, total = r.recv_views(recv_counts)
                c.all_to_all_v(r.send_views(counts), recv)
                next_local, deg_in = r.absorb(level, total)
                found, arcs, next_deg, sent = c.all_reduce_sum(
                    [next_local, arcs_l, deg_l + deg_in, sum(counts) - counts[c.rank]])
                self.levels.append(dict(direction="push", frontier=flen, arcs=arcs, discovered=found, sent=sent))
                r.swap()
                level += 1
                if found == 0:
                    break
                m_unexplored -= arcs
                if self.mode == "beamer" and next_deg > m_unexplored / self.alpha and found > flen:
                    direction = PULL
                    r.list_to_slice(next_local)
                flen_local, flen = next_local, found
            else:
                full, part = r.gather_buffers()
                c.all_gather_into(full, part)
                r.or_known()
                found_l, arcs_l, deg_l = r.pull(level)
                found, arcs = c.all_reduce_sum([found_l, arcs_l])
                self.levels.append(dict(direction="pull", frontier=flen, arcs=arcs, discovered=found, sent=0))
                level += 1
                if found == 0:
                    break
                if found < self.n / self.beta and found < flen:
                    # hand the (small) frontier back to push; peers' last pull discoveries become "known"
                    c.all_gather_into(full, part)
                    r.or_known()
                    flen_local = r.slice_to_list()
                    direction = PUSH
                flen = found
        return level


def verify_bfs_properties(rank_backend, comm, src: int):
    """Size-independent BFS checks at full scale, on the device with torch:
    labels[src] == 0; the reached set is closed under arcs; an arc spans at most one level; every
    reached vertex but the source has a neighbour one level up.  Returns (ok, level histogram)."""
    import torch
    r = rank_backend
    world, me = r.world, r.rank
    full = torch.empty(r.n_global, dtype=torch.int32, device=r.labels.device)
    comm.all_gather_into(full, r.labels)                     # rank-major: full[p * n_local + row]
    off = r.g.row_offsets.to(torch.int64) & 0xFFFFFFFF
    deg = off[1:] - off[:-1]
    ok = True
    log_p = PT.log2_ranks(world)
    if int(PT.owner(src, world)) == me:
        ok &= int(r.labels[src >> log_p].item()) == 0
    has_parent = torch.zeros(r.n_local, dtype=torch.bool, device=full.device)
    step = 1 << 22                                           # rows per chunk: bounds the temporaries
    for lo in range(0, r.n_local, step):
        hi = min(lo + step, r.n_local)
        a, b = int(off[lo].item()), int(off[hi].item())
        if a == b:
            continue
        rows = torch.repeat_interleave(torch.arange(lo, hi, device=full.device), deg[lo:hi])
        cols = r.g.col_indices[a:b].long()
        lv = r.labels[rows]
        crow = cols >> log_p                                  # b200::Partition::bit on the device
        cown = (cols & (world - 1)) ^ ((((crow * PT.GOLDEN) & 0xFFFFFFFF) >> (32 - log_p)) if log_p else 0)
        lu = full[cown * r.n_local + crow]
        ok &= bool(((lv >= 0) == (lu >= 0)).all())
        reached = lv >= 0
        if bool(reached.any()):
            ok &= int((lv[reached] - lu[reached]).abs().max().item()) <= 1
        has_parent[rows[reached & (lu == lv - 1)]] = True
    mine = r.labels
    need = mine > 0
    ok &= bool((has_parent | ~need).all())
    hist = torch.bincount((mine[mine >= 0]).long(), minlength=64)[:64]
    tot = comm.all_reduce_sum([int(ok)] + hist.tolist())
    hist_g = tot[1:]
    while hist_g and hist_g[-1] == 0:
        hist_g.pop()
    return tot[0] == world, hist_g
