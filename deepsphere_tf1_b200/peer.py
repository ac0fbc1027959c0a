"""One-shot all-reduce of the small SyncBN vectors over NVLink peer memory (csrc/peer_allreduce.cu).

Set-up (once per process group): every rank creates a communication buffer with `dsb200_peer_comm_create`, the
64-byte CUDA IPC handles travel through `torch.distributed.all_gather_object`, every rank opens its peers' buffers.
Afterwards `PeerAllReduce.allreduce(t)` is a single kernel launch on the current stream (CUDA-graph capturable).
torch.distributed stays the transport of the bulky fp32 gradient all-reduce and of the set-up itself."""
import ctypes
import os

import torch
import torch.distributed as dist

from ._lib import LIB

MAX_DOUBLES = 1024


class PeerAllReduce(object):
    def __init__(self, group=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError('peer all-reduce: at most 8 ranks (one NVSwitch node)')
        dll = LIB.load()
        dll.dsb200_peer_comm_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]
        dll.dsb200_peer_comm_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        dll.dsb200_peer_allreduce_f64.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_void_p]
        mine = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        rc = dll.dsb200_peer_comm_create(ctypes.byref(mine), handle)
        ok = torch.tensor([1 if rc == 0 else 0], device='cuda')
        handles = [None] * self.world
        dist.all_gather_object(handles, (handle.raw, os.uname().nodename), group=group)
        same_node = all(h[1] == handles[0][1] for h in handles)
        ptrs = (ctypes.c_void_p * self.world)()
        if rc == 0 and same_node:
            for r in range(self.world):
                if r == self.rank:
                    ptrs[r] = mine
                    continue
                p = ctypes.c_void_p()
                if dll.dsb200_peer_comm_open(handles[r][0], ctypes.byref(p)) != 0:
                    ok.zero_()
                    break
                ptrs[r] = p
        else:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)           # all ranks agree, or none uses it
        self.enabled = bool(int(ok.item()))
        self._dll, self._ptrs = dll, ptrs
        dist.barrier(group=group)

    def allreduce(self, t):
        """In-place sum over the ranks of a contiguous float64 CUDA tensor of at most 1024 elements."""
        n = t.numel()
        if not (self.enabled and t.dtype == torch.float64 and t.is_cuda and t.is_contiguous() and 0 < n <= MAX_DOUBLES):
            return False
        rc = self._dll.dsb200_peer_allreduce_f64(self._ptrs, self.rank, self.world, t.data_ptr(), n,
                                                 torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError('dsb200_peer_allreduce_f64 failed: ' + LIB.last_error())
        return True


class PeerHalo(object):
    """Mailboxes of the one-kernel halo exchange (csrc/peer_halo.cu): `capacity` floats per half on every rank, opened by all
    peers through CUDA IPC.  `exchange(sg, rows)` fills the ghost rows of `rows` [Me, F] in place on the current stream."""

    def __init__(self, capacity, group=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError('peer halo exchange: at most 8 ranks (one NVSwitch node)')
        dll = LIB.load()
        dll.dsb200_peer_halo_create.argtypes = [ctypes.c_longlong, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]
        dll.dsb200_peer_comm_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        dll.dsb200_peer_halo_exchange.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_void_p]
        cap = torch.tensor([(int(capacity) + 3) // 4 * 4], device='cuda', dtype=torch.int64)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
        self.capacity = int(cap.item())
        mine = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        rc = dll.dsb200_peer_halo_create(self.capacity, ctypes.byref(mine), handle)
        ok = torch.tensor([1 if rc == 0 else 0], device='cuda')
        handles = [None] * self.world
        dist.all_gather_object(handles, (handle.raw, os.uname().nodename), group=group)
        ptrs = (ctypes.c_void_p * self.world)()
        if rc == 0 and all(h[1] == handles[0][1] for h in handles):
            for r in range(self.world):
                if r == self.rank:
                    ptrs[r] = mine
                    continue
                p = ctypes.c_void_p()
                if dll.dsb200_peer_comm_open(handles[r][0], ctypes.byref(p)) != 0:
                    ok.zero_()
                    break
                ptrs[r] = p
        else:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        self.enabled = bool(int(ok.item()))
        self._dll, self._ptrs = dll, ptrs
        dist.barrier(group=group)

    def exchange(self, sg, rows):
        """rows: contiguous fp32 CUDA tensor viewed as [Me, F]; returns False when the shape is not covered (NCCL path)."""
        F = rows.shape[1]
        if not (self.enabled and rows.is_cuda and rows.dtype == torch.float32 and sg.H * F <= self.capacity):
            return False
        rc = self._dll.dsb200_peer_halo_exchange(self._ptrs, self.rank, self.world, rows.data_ptr(), F, sg.Ml, sg.H,
                                                 sg.halo_meta.data_ptr(), sg.halo_idx.data_ptr(), sg.halo_max_rows,
                                                 torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError('dsb200_peer_halo_exchange failed: ' + LIB.last_error())
        return True
