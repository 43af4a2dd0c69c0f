"""Domain decomposition and the NCCL communicator (Python side of the fvk_decomp_* / fvk_comm_* C ABI).

One process per GPU: every rank builds (or loads) the global mesh description on the host, cuts out its own sub-mesh
with ghost cells (`Decomposition`), uploads it (`UnstructuredMesh(dec.desc)`) and attaches the halo plan to a `Comm`.
torch.distributed is only used to hand the NCCL unique id from rank 0 to the other ranks."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, lib
from .mesh import MeshDesc


def simple_map(gdesc: MeshDesc, n) -> np.ndarray:
    out = np.zeros(gdesc.nCells, dtype=np.int32)
    check(lib().fvk_decomp_simple_map(C.byref(gdesc.c), C.c_int(n[0]), C.c_int(n[1]), C.c_int(n[2]), out.ctypes.data_as(C.c_void_p)))
    return out


def read_cell_decomposition(path, nCells: int) -> np.ndarray:
    """cell -> rank map from an OpenFOAM labelList file (`decomposePar -cellDist` writes constant/cellDecomposition);
    pass it as `cellRank=` to Decomposition (any decomposition method)."""
    out = np.zeros(nCells, dtype=np.int32)
    check(lib().fvk_labellist_read(str(path).encode(), C.c_int32(nCells), out.ctypes.data_as(C.c_void_p)))
    return out


def default_split(nRanks: int):
    """n (px py pz) for 1/2/4/8 ranks: 8 -> 2x2x2, 4 -> 2x2x1, 2 -> 2x1x1."""
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(nRanks) or (nRanks, 1, 1)


class Decomposition:
    def __init__(self, gdesc: MeshDesc, nRanks: int, rank: int, n=None, cellRank=None):
        self.nRanks, self.rank = nRanks, rank
        if cellRank is None:
            cellRank = simple_map(gdesc, n or default_split(nRanks))
        self.cellRank = np.ascontiguousarray(cellRank, dtype=np.int32)
        h = C.c_void_p()
        check(lib().fvk_decompose(C.byref(gdesc.c), self.cellRank.ctypes.data_as(C.c_void_p), C.c_int(nRanks), C.c_int(rank), C.byref(h)))
        self._h = h
        lib().fvk_decomp_mesh.restype = C.POINTER(_capi.MeshDesc)
        d = MeshDesc()
        d._c = lib().fvk_decomp_mesh(h).contents
        d._keep["decomp"] = self  # the arrays live in the fvk_decomp
        d.patch_names = list(gdesc.patch_names)
        self.desc = d
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().fvk_decomp_info(h, C.byref(a), C.byref(b), C.byref(c)))
        self.nOwned, self.nGhost, self.nNeighbours = a.value, b.value, c.value
        pc, pf = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        check(lib().fvk_decomp_maps(h, C.byref(pc), C.byref(pf)))
        self.cellGlobal = np.ctypeslib.as_array(pc, shape=(self.nOwned + self.nGhost,)).copy()
        self.faceGlobal = np.ctypeslib.as_array(pf, shape=(d.nFaces,)).copy() if d.nFaces else np.zeros(0, np.int32)
        pr, pso, psc, pro = (C.POINTER(C.c_int32)() for _ in range(4))
        check(lib().fvk_decomp_halo(h, C.byref(pr), C.byref(pso), C.byref(psc), C.byref(pro)))
        k = self.nNeighbours
        arr = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        self.nbrRanks = arr(pr, k)
        self.sendOff = arr(pso, k + 1) if k else np.zeros(1, np.int32)
        self.recvOff = arr(pro, k + 1) if k else np.zeros(1, np.int32)
        self.sendCells = arr(psc, int(self.sendOff[-1]))

    def __del__(self):
        if getattr(self, "_h", None) is not None and _capi is not None and _capi._lib is not None:
            _capi._lib.fvk_decomp_destroy(self._h)
            self._h = None

    # host-side scatter/gather helpers (tests, IO)
    def scatter_cells(self, g: np.ndarray) -> np.ndarray:
        """global cell field -> local field incl. ghost values"""
        return np.ascontiguousarray(g[self.cellGlobal])

    def scatter_faces(self, g: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(g[self.faceGlobal])

    def scatter_boundary(self, g: np.ndarray, nI_global: int) -> np.ndarray:
        nIl = self.desc.nInternalFaces
        return np.ascontiguousarray(g[self.faceGlobal[nIl:] - nI_global])


class Comm:
    """fvk_comm: NCCL communicator + halo plan. `Comm.from_torch()` takes rank/world from torch.distributed."""

    def __init__(self, rank: int, nRanks: int, unique_id: bytes | None):
        h = C.c_void_p()
        buf = (C.c_char * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        check(lib().fvk_comm_create(C.c_int(rank), C.c_int(nRanks), buf, C.byref(h)))
        self._h, self.rank, self.nRanks = h, rank, nRanks

    @classmethod
    def from_torch(cls):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        if world == 1:
            return cls(0, 1, None)
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            check(lib().fvk_comm_unique_id(buf))
            idt = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        idt = idt.to(dev)
        dist.broadcast(idt, 0)
        return cls(rank, world, bytes(idt.cpu().numpy().tobytes()))

    @property
    def handle(self):
        return self._h

    def set_halo(self, dec: Decomposition, p2p: bool | None = None):
        """Attach the halo plan of `dec`. p2p=True connects the peer-memory windows for it (collective over all ranks;
        None keeps the transport chosen by the last call)."""
        check(lib().fvk_comm_set_halo_from_decomp(self._h, dec._h))
        if p2p is not None:
            self._want_p2p = bool(p2p)
        if getattr(self, "_want_p2p", False) and self.nRanks > 1:
            self.enable_p2p()

    P2P_BLOB = 512

    def enable_p2p(self):
        """fvk_comm_p2p_export -> all-gather of the blobs over torch.distributed -> fvk_comm_p2p_connect. Collective.
        Returns False (and leaves NCCL as the transport on EVERY rank) when CUDA IPC is not available somewhere."""
        import torch
        import torch.distributed as dist
        blob = (C.c_char * self.P2P_BLOB)()
        check(lib().fvk_comm_p2p_export(self._h, blob))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        mine = torch.frombuffer(bytearray(blob.raw), dtype=torch.uint8).clone().to(dev)
        allb = torch.empty(self.nRanks * self.P2P_BLOB, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, mine)
        raw = bytes(allb.cpu().numpy().tobytes())
        rc = lib().fvk_comm_p2p_connect(self._h, (C.c_char * len(raw)).from_buffer_copy(raw))
        # the transport must be the same everywhere: if any rank could not map a window, all go back to NCCL
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            import sys
            if rc != 0:
                print(f"[fvk] rank {self.rank}: peer-memory windows unavailable ({lib().fvk_last_error().decode()}); using NCCL", file=sys.stderr)
            check(lib().fvk_comm_p2p_disable(self._h))
            return False
        dist.barrier()  # nobody pushes into a window before every rank has mapped and zeroed its own
        return True

    @property
    def p2p(self) -> bool:
        return bool(lib().fvk_comm_p2p_enabled(self._h))

    def halo_exchange(self, field):
        import torch
        ncomp = 3 if field.ndim == 2 else 1
        check(lib().fvk_comm_halo_exchange(self._h, C.c_void_p(field.data_ptr()), C.c_int(ncomp),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def halo_exchange_multi(self, fields):
        """one exchange for several cell fields (fvk_comm_halo_exchange_multi)"""
        import torch

        class _HF(C.Structure):
            _fields_ = [("field", C.c_void_p), ("ncomp", C.c_int32)]
        arr = (_HF * len(fields))(*[_HF(f.data_ptr(), 3 if f.ndim == 2 else 1) for f in fields])
        check(lib().fvk_comm_halo_exchange_multi(self._h, C.c_int(len(fields)), arr, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def p2p_debug(self):
        """fvk_comm_p2p_debug: accumulated ns / counts of the in-kernel communication phases (diagnostics)"""
        out = (C.c_uint64 * 8)()
        check(lib().fvk_comm_p2p_debug(self._h, out))
        return list(out)

    def allreduce_sum(self, t):
        import torch
        check(lib().fvk_comm_allreduce_sum(self._h, C.c_void_p(t.data_ptr()), C.c_int(t.numel()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def allreduce_max(self, t):
        import torch
        check(lib().fvk_comm_allreduce_max(self._h, C.c_void_p(t.data_ptr()), C.c_int(t.numel()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def close(self):
        if self._h is not None:
            lib().fvk_comm_destroy(self._h)
            self._h = None


def bind_host_to_device(device_index: int):
    """Pin this process to the CPUs NVML reports as local to CUDA device `device_index` (its PCIe root / NUMA node), so that the
    pinned host buffers allocated afterwards are first-touched on that node and host<->device copies do not cross the socket
    interconnect. One process per GPU: call it right after choosing the device. Returns the CPU set used, or None when NVML is
    unavailable, reports nothing, or the set does not intersect the CPUs this process may use (containers): nothing changes then."""
    import os

    import torch
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n = (os.cpu_count() or 1024) // 64 + 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use or use == allowed:
            return None if not use else sorted(use)
        os.sched_setaffinity(0, use)
        return sorted(use)
    except Exception:
        return None

