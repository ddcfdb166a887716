"""Peer-mapped device buffers for the row-sharded path (one process per GPU, one NVSwitch box).

`PeerBuffer(t, group)` exports the torch CUDA tensor `t` through CUDA IPC (csrc/hia_peer.cu),
exchanges the 64-byte handles with `torch.distributed` (plumbing) and maps every other rank's buffer
into this process: `ptr(r)` is a raw device pointer to rank r's copy, usable as a kernel argument
(P2P loads / stores over NVLink) or as the source of a copy-engine pull (`pull`). No tensor of a
peer is ever wrapped in torch -- the pointers only go to the C ABI.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class PeerBuffer:
    def __init__(self, t, group=None):
        assert t.is_cuda and t.is_contiguous()
        self.lib = _lib.load()
        self.t = t                                    # keeps the exported allocation alive
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._bases = {}
        self._ptrs = [None] * self.world
        self._ptrs[self.rank] = t.data_ptr()
        if self.world == 1:
            return
        handle = ctypes.create_string_buffer(64)
        off = ctypes.c_int64(0)
        _lib.check(self.lib.hia_ipc_export(t.data_ptr(), ctypes.addressof(handle), ctypes.addressof(off)),
                   "hia_ipc_export")
        mine = (bytes(handle.raw), int(off.value), t.numel() * t.element_size())
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=group)
        for r, (h, o, nbytes) in enumerate(allh):
            if r == self.rank:
                continue
            assert nbytes == mine[2], "peer buffers must have the same size on every rank"
            hb = ctypes.create_string_buffer(h, 64)
            base = ctypes.c_void_p(0)
            _lib.check(self.lib.hia_ipc_open(ctypes.addressof(hb), ctypes.addressof(base)), "hia_ipc_open")
            self._bases[r] = base.value
            self._ptrs[r] = base.value + o

    def ptr(self, r):
        """Device pointer (valid in this process) to rank r's buffer."""
        return self._ptrs[r]

    def pull(self, dst, r, src_byte_offset, nbytes, stream):
        """dst (local CUDA tensor) <- nbytes of rank r's buffer from src_byte_offset, on `stream`
        (a copy engine moves the data; no kernel is launched)."""
        assert dst.is_contiguous() and nbytes <= dst.numel() * dst.element_size()
        _lib.check(self.lib.hia_peer_copy(dst.data_ptr(), self._ptrs[r] + src_byte_offset, nbytes, stream),
                   "hia_peer_copy")

    def close(self):
        for r, base in list(self._bases.items()):
            self.lib.hia_ipc_close(base)
            del self._bases[r]

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
