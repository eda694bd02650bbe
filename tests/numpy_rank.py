"""NumPy stand-in for mini_b200.dist.GpuRank (TEST INFRASTRUCTURE, CPU only): implements the per-rank
steps of the partitioned BFS with the same partition rules (mini_b200.partition: swizzled-cyclic owner, row = v >> log P,
rank-major bitmaps) so the host-side control flow and the exchange protocol of DistBFS can run on gloo."""
import numpy as np
import torch

from mini_b200 import partition as PT


class NumpyRank:
    def __init__(self, csr, rank, world):
        self.rank, self.world = rank, world
        self.n, self.n_local = csr.n, csr.n // world
        rows = PT.global_ids(rank, world, csr.n // world)     # global id of every local row
        self.gid = rows
        self.off = np.concatenate([[0], np.cumsum(csr.offsets[rows + 1] - csr.offsets[rows])])
        self.idx = np.concatenate([csr.indices[csr.offsets[v]:csr.offsets[v + 1]] for v in rows]) if len(rows) else np.zeros(0, np.int32)
        self.labels = np.full(self.n_local, -1, np.int32)
        self.known = np.zeros(self.n, bool)            # indexed by rank-major bit position
        self.frontier_bits = torch.zeros(self.n, dtype=torch.uint8)
        self.next_slice = torch.zeros(self.n_local, dtype=torch.uint8)
        self.frontier = np.zeros(0, np.int32)
        self.next = []
        self.boxes = None
        self.inbox = torch.zeros(0, dtype=torch.int32)

    def bit(self, v):
        v = np.asarray(v, np.int64)
        return PT.bit(v, self.world, self.n_local).astype(np.int64)

    def owner(self, v):
        return int(PT.owner(int(v), self.world))

    def row(self, v):
        return int(v) >> PT.log2_ranks(self.world)

    def init(self, src):
        self.labels[:] = -1
        self.known[:] = False
        self.known[self.bit(src)] = True
        if self.owner(src) == self.rank:
            self.labels[self.row(src)] = 0
            self.frontier = np.array([src], np.int32)
            return 1
        self.frontier = np.zeros(0, np.int32)
        return 0

    def push(self, level, flen):
        boxes = [[] for _ in range(self.world)]
        arcs = deg = 0
        for v in self.frontier[:flen]:
            r = self.row(v)
            for u in self.idx[self.off[r]:self.off[r + 1]]:
                arcs += 1
                b = self.bit(u)
                if not self.known[b]:
                    self.known[b] = True
                    p = self.owner(u)
                    boxes[p].append(u)
                    if p == self.rank:
                        self.labels[self.row(u)] = level + 1
                        deg += self.off[self.row(u) + 1] - self.off[self.row(u)]
        self.next = boxes[self.rank]
        self.boxes = [torch.tensor(b, dtype=torch.int32) for b in boxes]
        return [len(b) for b in boxes], arcs, int(deg)

    def send_views(self, counts):
        return [self.boxes[p] if p != self.rank else torch.zeros(0, dtype=torch.int32) for p in range(self.world)]

    def recv_views(self, counts):
        total = sum(c for p, c in enumerate(counts) if p != self.rank)
        self.inbox = torch.zeros(total, dtype=torch.int32)
        out, off = [], 0
        for p in range(self.world):
            k = 0 if p == self.rank else counts[p]
            out.append(self.inbox[off:off + k])
            off += k
        return out, total

    def absorb(self, level, total):
        deg = 0
        for u in self.inbox[:total].numpy():
            b = self.bit(u)
            if not self.known[b]:
                self.known[b] = True
                self.labels[self.row(u)] = level + 1
                self.next.append(u)
                deg += self.off[self.row(u) + 1] - self.off[self.row(u)]
        return len(self.next), int(deg)

    def swap(self):
        self.frontier = np.array(self.next, np.int32)

    def list_to_slice(self, flen):
        s = np.zeros(self.n_local, np.uint8)
        s[np.asarray(self.frontier[:flen], np.int64) >> PT.log2_ranks(self.world)] = 1
        self.next_slice = torch.from_numpy(s)

    def gather_buffers(self):
        return self.frontier_bits, self.next_slice

    def or_known(self):
        self.known |= self.frontier_bits.numpy().astype(bool)

    def pull(self, level):
        fb = self.frontier_bits.numpy()
        nxt = np.zeros(self.n_local, np.uint8)
        found = arcs = 0
        base = self.rank * self.n_local
        for r in range(self.n_local):
            if self.known[base + r]:
                continue
            for u in self.idx[self.off[r]:self.off[r + 1]]:
                arcs += 1
                if fb[self.bit(u)]:
                    self.labels[r] = level + 1
                    self.known[base + r] = True
                    nxt[r] = 1
                    found += 1
                    break
        self.next_slice = torch.from_numpy(nxt)
        return found, arcs, 0

    def slice_to_list(self):
        rows = np.flatnonzero(self.next_slice.numpy())
        self.frontier = self.gid[rows].astype(np.int32)
        self.next = list(self.frontier)
        return len(rows)
