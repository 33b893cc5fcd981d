"""bosql_b200.distributed — the host side of the multi-GPU operator layer (one process per GPU, torch.distributed).

The C++ operators decide WHAT crosses NVLink (partial aggregate states, join bitmaps, broadcast build sides; the key-hash
shuffle stores straight into peer memory and needs only host-side size / handle exchanges) and call a five-entry C function
table at those points (include/bosql_b200_exec.h: bqx_exchange).  This module fills that table with torch.distributed calls:
NCCL on device pointers, gloo on host pointers (CPU tests), device pointers staged over gloo (several processes sharing one
GPU, which NCCL refuses).  `install()` makes the current process one rank of the job; nothing else is needed - the same SQL
statement through `Engine.plan(...).run()` then runs sharded.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _gather_cpu(g, pad, world, group):
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    torch.cat(parts, out=g)


class _CudaArray:
    """__cuda_array_interface__ over a raw device pointer, so torch can view device memory it did not allocate."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


# ---- the exchange table (include/bosql_b200_exec.h: bqx_exchange) ------------------------------------------------------
import ctypes as _C

_I64P = _C.POINTER(_C.c_int64)
_CB_GATHER = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_size_t, _C.c_void_p)
_CB_GATHER_V = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.c_void_p, _C.c_void_p, _I64P, _C.c_void_p)
_CB_A2A_V = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.c_void_p, _I64P, _C.c_void_p, _I64P, _C.c_void_p)
_CB_SUM = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _C.c_void_p, _C.c_size_t, _C.c_void_p)
_CB_HOST = _C.CFUNCTYPE(_C.c_int, _C.c_void_p, _I64P, _C.c_int32, _I64P)


class ExchangeTable(_C.Structure):
    _fields_ = [("user", _C.c_void_p), ("world", _C.c_int32), ("rank", _C.c_int32), ("keep_sharded", _C.c_int32),
                ("pad", _C.c_int32), ("all_gather", _CB_GATHER), ("all_gather_v", _CB_GATHER_V), ("all_to_all_v", _CB_A2A_V),
                ("all_reduce_sum_u32", _CB_SUM), ("host_all_gather_i64", _CB_HOST)]


class Exchange:
    """Collectives for the multi-GPU operator layer.  device="cuda": pointers are device memory, ordered on the stream the
    operators pass (NCCL).  device="cpu": pointers are host memory (gloo) - the same code path, used by the CPU tests.
    device="cuda" under a gloo group: device pointers staged through host tensors - slow, but it lets two processes share
    ONE GPU (NCCL refuses that), which is how the operator-layer exchange is tested on a single-GPU box."""

    def __init__(self, group=None, device="cuda", keep_sharded=False):
        self.group = group
        self.device = device
        self.staged = device == "cuda" and dist.get_backend(group) == "gloo"
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._views = {}             # (ptr, nbytes) -> uint8 tensor view (non-owning; building one costs ~20 us)
        self._streams = {}           # cudaStream_t -> torch.cuda.ExternalStream
        self.calls = {"all_gather": 0, "all_gather_v": 0, "all_to_all_v": 0, "all_reduce_sum_u32": 0, "host_all_gather_i64": 0}
        self.bytes_sent = 0
        self.error = None
        self._cbs = (_CB_GATHER(self._all_gather), _CB_GATHER_V(self._all_gather_v), _CB_A2A_V(self._all_to_all_v),
                     _CB_SUM(self._sum_u32), _CB_HOST(self._host_gather))
        self.table = ExchangeTable(None, self.world, self.rank, 1 if keep_sharded else 0, 0, *self._cbs)

    # -- views over raw pointers ------------------------------------------------------------------------------------
    def _bytes(self, ptr, n):
        if n == 0:
            return torch.empty(0, dtype=torch.uint8, device=self.device)
        if self.device == "cpu":
            return torch.frombuffer((_C.c_char * n).from_address(ptr), dtype=torch.uint8)
        key = (ptr, n)
        t = self._views.get(key)
        if t is None:
            if len(self._views) > 256:
                self._views.clear()
            t = self._views[key] = torch.as_tensor(_CudaArray(ptr, n, "|u1"), device="cuda")
        return t

    def _on(self, stream):
        if self.device == "cpu":
            import contextlib
            return contextlib.nullcontext()
        st = self._streams.get(stream)
        if st is None:
            st = self._streams[stream] = torch.cuda.ExternalStream(stream or 0)
        return torch.cuda.stream(st)

    def _guard(self, name, fn):
        try:
            self.calls[name] += 1
            fn()
            return 0
        except Exception as e:  # noqa: BLE001 - reported through the C status code
            self.error = f"{name}: {e}"
            return 1

    # -- callbacks --------------------------------------------------------------------------------------------------
    def _all_gather(self, user, send, recv, nbytes, stream):
        def go():
            with self._on(stream):
                src, dst = self._bytes(send, nbytes), self._bytes(recv, nbytes * self.world)
                self.bytes_sent += nbytes
                if self.staged:
                    h = torch.empty(dst.numel(), dtype=torch.uint8)
                    _gather_cpu(h, src.cpu(), self.world, self.group)
                    dst.copy_(h)
                elif self.device == "cpu":
                    _gather_cpu(dst, src, self.world, self.group)
                else:
                    dist.all_gather_into_tensor(dst, src, group=self.group)
        return self._guard("all_gather", go)

    def _all_gather_v(self, user, send, recv, bytes_by_rank, stream):
        def go():
            sizes = [int(bytes_by_rank[r]) for r in range(self.world)]
            with self._on(stream):
                dst = self._bytes(recv, sum(sizes))
                src = self._bytes(send, sizes[self.rank])
                self.bytes_sent += sizes[self.rank]
                device_dst = None
                if self.staged:
                    device_dst, dst, src = dst, torch.empty(dst.numel(), dtype=torch.uint8), src.cpu()
                if not self.staged and self.device == "cuda" and len(set(sizes)) == 1 and sizes[0]:
                    dist.all_gather_into_tensor(dst, src, group=self.group)        # equal shards: one collective
                    return
                off = 0
                # one broadcast per contributing rank, straight into its slot of the output (no padding, no staging copy)
                for r, n in enumerate(sizes):
                    if n:
                        slot = dst[off:off + n]
                        if r == self.rank:
                            slot.copy_(src)
                        dist.broadcast(slot, src=dist.get_global_rank(self.group, r) if self.group is not None else r, group=self.group)
                    off += n
                if device_dst is not None:
                    device_dst.copy_(dst)
        return self._guard("all_gather_v", go)

    def _all_to_all_v(self, user, send, send_bytes, recv, recv_bytes, stream):
        def go():
            sb = [int(send_bytes[r]) for r in range(self.world)]
            rb = [int(recv_bytes[r]) for r in range(self.world)]
            with self._on(stream):
                src, dst = self._bytes(send, sum(sb)), self._bytes(recv, sum(rb))
                self.bytes_sent += sum(sb) - sb[self.rank]
                device_dst = None
                if self.staged:
                    device_dst, dst, src = dst, torch.empty(dst.numel(), dtype=torch.uint8), src.cpu()
                if self.device == "cpu" or self.staged:
                    outs = list(dst.split(rb)) if sum(rb) else [torch.empty(0, dtype=torch.uint8) for _ in rb]
                    ins = list(src.split(sb)) if sum(sb) else [torch.empty(0, dtype=torch.uint8) for _ in sb]
                    _all_to_all_cpu(outs, ins, self.rank, self.world, self.group)
                    if device_dst is not None:
                        device_dst.copy_(dst)
                else:
                    dist.all_to_all_single(dst, src, output_split_sizes=rb, input_split_sizes=sb, group=self.group)
        return self._guard("all_to_all_v", go)

    def _sum_u32(self, user, buf, words, stream):
        def go():
            with self._on(stream):
                t = self._bytes(buf, words * 4).view(torch.int32)      # two's-complement sum == unsigned sum, bit for bit
                self.bytes_sent += words * 4
                if self.staged:
                    h = t.cpu()
                    dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.group)
                    t.copy_(h)
                else:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return self._guard("all_reduce_sum_u32", go)

    def _host_gather(self, user, mine, n, out):
        def go():
            where = "cpu" if self.staged else self.device
            t = torch.tensor([int(mine[i]) for i in range(n)], dtype=torch.int64, device=where)
            g = torch.empty(n * self.world, dtype=torch.int64, device=where)
            if where == "cpu":
                _gather_cpu(g, t, self.world, self.group)
            else:
                dist.all_gather_into_tensor(g, t, group=self.group)
            for i, v in enumerate(g.cpu().tolist()):
                out[i] = v
        return self._guard("host_all_gather_i64", go)

    # -- installation -----------------------------------------------------------------------------------------------
    def install(self, exec_library):
        """bqx_set_exchange(&table): from now on every plan in this process runs as one rank of `group`."""
        if exec_library.bqx_set_exchange(_C.byref(self.table)):
            raise RuntimeError(exec_library.bqx_last_error().decode())
        return self

    @staticmethod
    def uninstall(exec_library):
        exec_library.bqx_set_exchange(None)


def _all_to_all_cpu(outs, ins, rank, world, group):
    """gloo: pairwise isend / irecv (sizes differ per pair, which gloo's scatter does not allow)."""
    g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    outs[rank].copy_(ins[rank])
    reqs = []
    for peer in range(world):
        if peer == rank:
            continue
        if ins[peer].numel():
            reqs.append(dist.isend(ins[peer].contiguous(), dst=g(peer), group=group))
        if outs[peer].numel():
            reqs.append(dist.irecv(outs[peer], src=g(peer), group=group))
    for r in reqs:
        r.wait()


_INSTALLED = None


class NativeExchange:
    """The library's own NCCL exchange (bqx_comm_init): after this call no Python runs between plan.run() and a collective.
    torch.distributed is used ONCE, to hand rank 0's NCCL unique id to the other ranks."""

    def __init__(self, exec_library, group=None, keep_sharded=False):
        self.lib = exec_library
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        ident = (_C.c_ubyte * 128)()
        if self.rank == 0 and exec_library.bqx_comm_unique_id(ident):
            raise RuntimeError(exec_library.bqx_last_error().decode())
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (_C.c_ubyte * 128).from_buffer_copy(box[0])
        if exec_library.bqx_comm_init(self.world, self.rank, ident, 1 if keep_sharded else 0):
            raise RuntimeError(exec_library.bqx_last_error().decode())
        self.error = None

    KINDS = ("all_gather", "all_gather_v", "all_to_all_v", "all_reduce_sum_u32", "host_all_gather_i64")

    def _stats(self):
        calls, sent = (_C.c_uint64 * 5)(), _C.c_uint64(0)
        if self.lib.bqx_comm_stats(calls, _C.byref(sent)):
            raise RuntimeError(self.lib.bqx_last_error().decode())
        return [int(c) for c in calls], int(sent.value)

    @property
    def calls(self):
        return dict(zip(self.KINDS, self._stats()[0]))

    def keep_sharded(self, on):
        self.lib.bqx_exchange_keep_sharded(1 if on else 0)

    @property
    def bytes_sent(self):
        return self._stats()[1]


def install_native(exec_library=None, group=None, keep_sharded=False):
    """Make this process one rank of `group` with the collectives inside libbosql_b200.so (NCCL over NVLink)."""
    global _INSTALLED
    if exec_library is None:
        from .engine import exec_lib
        exec_library = exec_lib()
    _INSTALLED = NativeExchange(exec_library, group, keep_sharded)
    return _INSTALLED


def install(exec_library=None, group=None, device="cuda", keep_sharded=False):
    """Make this process one rank of `group` for every plan it runs from now on (tables hold row shards; statistics passed
    to Engine.add_table describe the whole table).  Returns the Exchange (its .calls / .bytes_sent count the traffic)."""
    global _INSTALLED
    if exec_library is None:
        from .engine import exec_lib
        exec_library = exec_lib()
    _INSTALLED = Exchange(group, device, keep_sharded).install(exec_library)
    return _INSTALLED


def uninstall(exec_library=None):
    global _INSTALLED
    if exec_library is None:
        from .engine import exec_lib
        exec_library = exec_lib()
    Exchange.uninstall(exec_library)
    _INSTALLED = None
