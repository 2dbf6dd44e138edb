"""Symmetric memory over NVLink 5 / NVSwitch: the substrate of the native collectives.

One allocation of identical size on every rank, peer-mapped into every process and — when the fabric
offers it — bound to an NVLS multicast object.  Only the *handles* travel through torch (its
``_symmetric_memory`` rendezvous does the cuMem export/import + multicast bind; if that is unavailable
we fall back to CUDA-IPC handles exchanged with ``all_gather_object``, which gives peer pointers but no
multicast).  Every byte of gradient / parameter traffic afterwards is moved by our own kernels
(``csrc/comm_sm100.cu``, ``csrc/gemm_sm100.cu``).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .. import ops

_FORCE_NO_MC = os.environ.get("TDS_NO_MULTICAST", "0") == "1"
_FORCE_IPC = os.environ.get("TDS_SYMM_IPC", "0") == "1"


def available() -> bool:
    if not (torch.cuda.is_available() and dist.is_available() and dist.is_initialized()):
        return False
    if dist.get_backend() != "nccl":
        return False
    return os.environ.get("TDS_DISABLE_NATIVE", "0") != "1"


class SymmTensor:
    """A flat symmetric buffer: ``local`` is this rank's tensor, ``peer(r)`` a tensor aliasing rank r's copy."""

    def __init__(self, local: torch.Tensor, peer_ptrs: List[int], mc_ptr: int, handle=None, keep=None):
        self.local = local
        self.peer_ptrs = peer_ptrs
        self.mc_ptr = 0 if _FORCE_NO_MC else int(mc_ptr or 0)
        self.handle = handle
        self._keep = keep            # IPC-opened storages must stay alive
        self.buf = ops.ext().SymmBuf([int(p) for p in peer_ptrs], self.mc_ptr)

    @property
    def has_multicast(self) -> bool:
        return self.mc_ptr != 0

    def peer(self, rank: int, shape, dtype, byte_offset: int = 0) -> torch.Tensor:
        """Tensor view of ``shape``/``dtype`` at ``byte_offset`` inside rank ``rank``'s copy."""
        esz = torch.empty(0, dtype=dtype).element_size()
        assert byte_offset % esz == 0
        if self.handle is not None:
            return self.handle.get_buffer(rank, list(shape), dtype, byte_offset // esz)
        st = self._keep[rank]
        n = 1
        for s in shape:
            n *= s
        return torch.empty(0, dtype=dtype, device=self.local.device).set_(st, byte_offset // esz, tuple(shape))


def _alloc_torch_symm(nbytes: int, device, group) -> Optional[SymmTensor]:
    try:
        import torch.distributed._symmetric_memory as sm
        t = sm.empty(nbytes, dtype=torch.uint8, device=device)
        hdl = sm.rendezvous(t, group=group.group_name if group is not None else dist.group.WORLD.group_name)
        mc = 0
        try:
            if hdl.has_multicast_support:
                mc = int(hdl.multicast_ptr)
        except Exception:
            mc = 0
        return SymmTensor(t, [int(p) for p in hdl.buffer_ptrs], mc, handle=hdl)
    except Exception as e:  # pragma: no cover - depends on the box
        if os.environ.get("TDS_DEBUG"):
            print(f"[tds symm] torch symmetric memory unavailable: {type(e).__name__}: {e}")
        return None


def _alloc_ipc(nbytes: int, device, group) -> SymmTensor:
    """Peer-mapped buffers through CUDA IPC (no multicast)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    # a dedicated cudaMalloc block (IPC shares whole allocations): bypass the caching allocator's pooling
    t = torch.empty(nbytes, dtype=torch.uint8, device=device)
    st = t.untyped_storage()
    share = st._share_cuda_()
    gathered = [None] * world
    dist.all_gather_object(gathered, share, group=group)
    storages, ptrs = [], []
    for r in range(world):
        if r == rank:
            storages.append(st)
            ptrs.append(t.data_ptr())
        else:
            s = torch.UntypedStorage._new_shared_cuda(*gathered[r])
            storages.append(s)
            ptrs.append(s.data_ptr())
    return SymmTensor(t, ptrs, 0, handle=None, keep=storages)


def alloc(nbytes: int, device, group=None) -> SymmTensor:
    """Collective: every rank calls with the same size.  Memory is zero-initialised."""
    nbytes = (int(nbytes) + 4095) // 4096 * 4096
    st = None if _FORCE_IPC else _alloc_torch_symm(nbytes, device, group)
    # all ranks must agree on the path
    ok = torch.tensor([1 if st is not None else 0], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        st = _alloc_ipc(nbytes, device, group)
    st.local.zero_()
    torch.cuda.synchronize(device)
    dist.barrier(group=group)
    return st


class Comm:
    """Flag pads + context for the device-side barriers of our collectives."""

    def __init__(self, device, group=None):
        ext = ops.ext()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= ext.COMM_MAX_RANKS
        words = 8 * ext.COMM_MAX_BLOCKS * ext.COMM_MAX_RANKS      # 8 channels
        self.flags = alloc(words * 4, device, group)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        self.ctx = ext.CommCtx([int(p) for p in self.flags.peer_ptrs], self.rank, self.world, self.error)
        self.device = device

    def barrier(self, channel: int = 7):
        ops.ext().comm_barrier(self.ctx, channel)
        ops.count_launch()

    def allreduce(self, st: SymmTensor, elem_off: int, numel: int, *, f32=False, scale=1.0, blocks=32, channel=0):
        ops.ext().comm_allreduce(self.ctx, st.buf, int(elem_off), int(numel), bool(f32), float(scale), int(blocks), int(channel))
        ops.count_launch()

    def reduce_to(self, st: SymmTensor, elem_off: int, numel: int, dst: int, *, f32=False, scale=1.0, blocks=32, channel=0):
        ops.ext().comm_reduce_to(self.ctx, st.buf, int(elem_off), int(numel), bool(f32), int(dst), float(scale), int(blocks), int(channel))
        ops.count_launch()

    def broadcast(self, st: SymmTensor, byte_off: int, nbytes: int, src: int, *, blocks=32, channel=0):
        ops.ext().comm_broadcast(self.ctx, st.buf, int(byte_off), int(nbytes), int(src), int(blocks), int(channel))
        ops.count_launch()
