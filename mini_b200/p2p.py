"""Peer-memory multi-GPU BFS (include/b200_frontier.h, b200_p2p_bfs_*): one rank per GPU, the frontier
exchange fused into the kernels over NVLink.  This module only sets the ranks up -- it allocates the
rank's symmetric heap, trades the CUDA IPC handles through torch.distributed (or plain addresses when
the ranks are threads of one process) and calls the C entry point that runs the whole traversal."""
from __future__ import annotations

import ctypes as C

from . import lib as L

IPC_HANDLE_BYTES = 64


class P2PBfs:
    def __init__(self, ctx, rank: int, world: int, n_global: int, m_global: int, graph_local):
        import torch
        self._L = L.load_library()
        self.ctx, self.rank, self.world = ctx, rank, world
        self.n_global, self.m_global, self.n_local = n_global, m_global, n_global // world
        self.g = graph_local
        self.cg = graph_local.cview()
        self._h = C.c_void_p()
        self._handle = (C.c_ubyte * IPC_HANDLE_BYTES)()
        base = C.c_void_p()
        L._check(self._L.b200_p2p_bfs_create(ctx._h, rank, world, n_global, C.byref(self._h), self._handle, C.byref(base)),
                 "b200_p2p_bfs_create")
        self.heap_base = base.value
        self.labels = torch.empty(self.n_local, dtype=torch.int32, device=ctx.torch_device)
        self.levels = []
        self.device_ms = 0.0
        self.launches = 0
        if world == 1:
            self.connect_local([self.heap_base])

    # -- wiring ------------------------------------------------------------------------------
    def ipc_handle(self) -> bytes:
        return bytes(self._handle)

    def connect_ipc(self, handles: bytes):
        """handles: world * 64 bytes, rank-major (every rank's ipc_handle())."""
        assert len(handles) == self.world * IPC_HANDLE_BYTES
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        L._check(self._L.b200_p2p_bfs_connect(self._h, buf, None), "b200_p2p_bfs_connect")

    def connect_local(self, bases):
        arr = (C.c_void_p * self.world)(*bases)
        L._check(self._L.b200_p2p_bfs_connect(self._h, None, arr), "b200_p2p_bfs_connect")

    def connect_torch_distributed(self):
        """All-gather the IPC handles over the default process group (NCCL or gloo) and map the peers."""
        import torch
        import torch.distributed as dist
        dev = self.ctx.torch_device if dist.get_backend() == "nccl" else torch.device("cpu")
        mine = torch.frombuffer(bytearray(self.ipc_handle()), dtype=torch.uint8).to(dev)
        rows = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(rows, mine)
        self.connect_ipc(b"".join(bytes(r.cpu().numpy().tobytes()) for r in rows))
        dist.barrier()

    # -- the traversal -----------------------------------------------------------------------
    def prepare(self, mode: str = "beamer"):
        """Builds the traversal graph ahead of the first run (call on every rank, then barrier)."""
        m = L.BFS_BEAMER if mode == "beamer" else L.BFS_PUSH
        L._check(self._L.b200_p2p_bfs_prepare(self._h, C.byref(self.cg), m, self.labels.data_ptr()), "b200_p2p_bfs_prepare")

    def run(self, src: int = 0, mode: str = "beamer", alpha: float = 15.0, beta: float = 18.0, timing: bool = False):
        cs = L.CStats()
        cs.collect_timing = int(timing)
        sent = (C.c_int64 * L.MAX_LEVELS)()
        m = L.BFS_BEAMER if mode == "beamer" else L.BFS_PUSH
        L._check(self._L.b200_p2p_bfs_run(self._h, C.byref(self.cg), self.m_global, src, m, alpha, beta,
                                          self.labels.data_ptr(), C.byref(cs), sent), "b200_p2p_bfs_run")
        st = L.Stats(cs)
        self.levels = [dict(direction=l["direction"], frontier=l["frontier_len"], arcs=l["arcs"], discovered=l["discovered"],
                            sent=int(sent[i]), level_ms=l["level_ms"], exchange=l["exchange"]) for i, l in enumerate(st.levels)]
        self.device_ms, self.launches, self.level_loop = st.device_ms, st.launches, st.level_loop
        return st.num_levels

    def set_trace(self, on: bool = True):
        L._check(self._L.b200_p2p_bfs_set_trace(self._h, int(on)), "b200_p2p_bfs_set_trace")

    def last_trace(self):
        """[(microseconds since the first entry, id)] of the last traced run (ids: include/b200_frontier.h)."""
        cap = 4096
        buf = (C.c_uint64 * cap)()
        n = C.c_int64()
        L._check(self._L.b200_p2p_bfs_last_trace(self._h, buf, cap, C.byref(n)), "b200_p2p_bfs_last_trace")
        if n.value == 0:
            return []
        t0 = buf[0] >> 8
        return [(((buf[i] >> 8) - t0) * 1e-3, int(buf[i] & 255)) for i in range(n.value)]

    def level_times_ms(self):
        """Per-level device time of the last traced run, from the "level closed" markers (64 + level)."""
        tr = self.last_trace()
        out, prev = [], 0.0
        for t, k in tr:
            if k >= 64:
                out.append((t - prev) * 1e-3)
                prev = t
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200_p2p_bfs_destroy(self._h)
            self._h = None

    # -- helpers shared with dist.GpuRank (verification, TEPS numerator) ------------------------
    def reached_degree_sum(self) -> int:
        import torch
        off = self.g.row_offsets.to(torch.int64) & 0xFFFFFFFF
        deg = off[1:] - off[:-1]
        return int(deg[self.labels >= 0].sum().item())

    def reached_count(self) -> int:
        return int((self.labels >= 0).sum().item())
