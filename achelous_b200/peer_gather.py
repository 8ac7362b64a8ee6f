"""All-gather of the packed output rows over NVLink peer memory, driven by copy engines.

One process per GPU (`torch.distributed` is only the plumbing that ships the IPC handles).  Every rank owns one exportable
allocation `[slot 0 | slot 1 | flag words]`; slot s is a (world * rows, width) matrix and the Engine of that slot writes this
rank's rows IN PLACE (`Engine(out=gather.slot(s)[gather.my_rows])`).  `push` then copies those rows into the same place of every
peer's slot with `cudaMemcpyAsync` between peer-mapped pointers - the transfers run on the copy engines, so no SM (and no NCCL
channel CTA) is taken from the forward kernels of the next step - and raises this rank's "ready" word on every peer.

Flag protocol (all values are the caller's monotonically increasing step number `seq` >= 1; nothing is ever reset):
    push(s, seq)     wait until every peer released what slot s held before (free[s][*] >= seq - slots), copy, ready[s][me] = seq on all
    wait(s, seq)     the stream continues once ready[s][*] >= seq: slot s holds the rows of all ranks for step seq
    release(s, seq)  free[s][me] = seq on all ranks: this rank has consumed slot s

Replaces the gather half of nn.DataParallel (/root/reference/achelous.py:176-177) for multi-process serving; the NCCL
all-gather it supersedes needed 16-32 channel CTAs for 2.07 GB per rank and step (profiles/r2_bench_8gpu.json)."""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class _RawCuda:
    """numba-style view of a raw device allocation, so that torch can wrap it without copying"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerGather:
    def __init__(self, rows, width, dtype, device, group=None, slots=2, copy_streams=2):
        self.lib = _lib.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert self.world <= 64, "ach_peer_signal / ach_peer_wait take at most 64 flag words"
        self.device = torch.device(device)
        self.rows, self.width, self.dtype, self.nslots = rows, width, dtype, slots
        item = torch.empty(0, dtype=dtype).element_size()
        self.row_bytes = width * item
        self.part_bytes = rows * self.row_bytes
        self.slot_bytes = (self.world * self.part_bytes + 255) // 256 * 256
        self.flags_off = slots * self.slot_bytes
        total = self.flags_off + 4096
        assert slots * 2 * self.world * 4 <= 4096
        with torch.cuda.device(self.device):
            p = C.c_void_p()
            _lib.check(self.lib.ach_peer_alloc(total, C.byref(p)), "ach_peer_alloc")
            self.base = p.value
            hb = self.lib.ach_peer_handle_bytes()
            buf = C.create_string_buffer(hb)
            _lib.check(self.lib.ach_peer_export(self.base, buf), "ach_peer_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, buf.raw, group=group)
            self.peer_base = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.peer_base.append(self.base)
                    continue
                q = C.c_void_p()
                _lib.check(self.lib.ach_peer_open(h, C.byref(q)), f"ach_peer_open(rank {r})")
                self.peer_base.append(q.value)
            self._raw = torch.as_tensor(_RawCuda(self.base, total), device=self.device)
            self._slots = [self._raw[s * self.slot_bytes: s * self.slot_bytes + self.world * self.part_bytes].view(dtype)
                           .view(self.world * rows, width) for s in range(slots)]
            # device arrays of flag addresses: [kind][slot] -> this rank's word in every rank's allocation
            addr = [[[b + self._flag(s, kind, self.rank) for b in self.peer_base] for s in range(slots)] for kind in (0, 1)]
            self._flag_ptrs = torch.tensor(addr, dtype=torch.int64, device=self.device)     # (2, slots, world)
            self._streams = [torch.cuda.Stream(self.device) for _ in range(max(1, copy_streams))]
        self.my_rows = slice(self.rank * rows, (self.rank + 1) * rows)
        # every rank has mapped every allocation before anybody pushes
        dist.barrier(group=group)

    def _flag(self, s, kind, r):   # byte offset of flag word (kind 0: ready, 1: free) of rank r for slot s
        return self.flags_off + ((s * 2 + kind) * self.world + r) * 4

    def slot(self, s):
        """(world * rows, width) tensor: the gathered rows of slot s (rank r's rows at [r * rows, (r + 1) * rows))"""
        return self._slots[s]

    def push(self, s, seq, stream):
        """After everything already on `stream`: send this rank's rows of slot s to every peer (copy engines), then mark them ready."""
        st = stream.cuda_stream
        if seq > self.nslots:
            _lib.check(self.lib.ach_peer_wait(self.base + self._flag(s, 1, 0), self.world, seq - self.nslots, st), "ach_peer_wait")
        off = s * self.slot_bytes + self.rank * self.part_bytes
        start = stream.record_event()
        used = []
        for k in range(1, self.world):
            p = (self.rank + k) % self.world          # rotated order: at any moment the ranks target different receivers
            cs = self._streams[(k - 1) % len(self._streams)]
            if cs not in used:
                cs.wait_event(start)
                used.append(cs)
            _lib.check(self.lib.ach_peer_copy(self.peer_base[p] + off, self.base + off, self.part_bytes, cs.cuda_stream), "ach_peer_copy")
        for cs in used:
            stream.wait_stream(cs)
        _lib.check(self.lib.ach_peer_signal(self._flag_ptrs[0, s].data_ptr(), self.world, seq, st), "ach_peer_signal")

    def wait(self, s, seq, stream):
        """`stream` continues once slot s holds the rows of ALL ranks for step seq."""
        _lib.check(self.lib.ach_peer_wait(self.base + self._flag(s, 0, 0), self.world, seq, stream.cuda_stream), "ach_peer_wait")

    def release(self, s, seq, stream):
        """After everything already on `stream`: tell every rank that this rank is done reading slot s of step seq."""
        _lib.check(self.lib.ach_peer_signal(self._flag_ptrs[1, s].data_ptr(), self.world, seq, stream.cuda_stream), "ach_peer_signal")

    def close(self):
        if getattr(self, "base", None) is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)              # nobody is still copying into a mapping that is about to go away
        self._slots, self._raw = None, None
        with torch.cuda.device(self.device):
            for r, b in enumerate(self.peer_base):
                if r != self.rank:
                    self.lib.ach_peer_close(b)
            dist.barrier(group=self.group)
            self.lib.ach_peer_free(self.base)
        self.base = None
