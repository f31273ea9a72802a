"""A plain ncclComm_t for the engine's in-library exchange step (lux_ddgi_set_nccl_comm), created through ctypes on the same
libnccl.so.2 the engine dlopens.  The unique id travels over an existing torch.distributed process group (any backend)."""
import ctypes as C


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


class NcclComm:
    def __init__(self, rank: int, world: int, device: int):
        import torch
        import torch.distributed as dist

        self._lib = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)  # resolves to the already loaded copy (torch's) when there is one
        self._lib.ncclGetErrorString.restype = C.c_char_p
        uid = _UniqueId()
        if rank == 0:
            self._check(self._lib.ncclGetUniqueId(C.byref(uid)))
        use_cuda = dist.get_backend() == "nccl"
        t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).clone()
        if use_cuda:
            t = t.cuda(device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
        C.memmove(C.byref(uid), raw, 128)
        torch.cuda.set_device(device)
        comm = C.c_void_p()
        self._lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        self._check(self._lib.ncclCommInitRank(C.byref(comm), world, uid, rank))
        self.ptr = comm.value

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("NCCL: " + self._lib.ncclGetErrorString(rc).decode())

    def destroy(self):
        if getattr(self, "ptr", None):
            self._lib.ncclCommDestroy.argtypes = [C.c_void_p]
            self._lib.ncclCommDestroy(C.c_void_p(self.ptr))
            self.ptr = None
