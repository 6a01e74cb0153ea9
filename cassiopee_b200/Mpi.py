"""Process-per-GPU plumbing for the distributed chimera path.

The reference's Converter.Mpi wraps mpi4py (rank, size, allgather, pickled
isend/recv following a {src:{dst:[...]}} graph — Converter/Converter/Mpi4py.py:149-168).
Here ranks are the processes started by torchrun (RANK / WORLD_SIZE /
MASTER_ADDR / MASTER_PORT), small Python objects travel through a
torch.distributed *gloo* group (plumbing only), and bulk data travels
  * device to device over NVLink through the library's own NCCL communicator
    (cx_exchange_dev / cx_exchange_bytes_dev), or
  * host to host through gloo when no GPU is present (CPU tests of the host
    logic; the compute backend is then injected by the test).
"""
import os
import numpy as np


class Comm:
    """Communicator.  device=True additionally opens the NCCL communicator on `ctx`."""

    def __init__(self, ctx=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.size = int(os.environ.get("WORLD_SIZE", "1"))
        self.ctx = ctx
        self.device = (ctx is not None) if device is None else device
        self._dist = None
        self._nccl_ready = False
        if self.size > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group(backend="gloo", rank=self.rank, world_size=self.size)
            self._dist = dist
        if self.device and self.size > 1:
            from . import _lib
            import ctypes
            idbuf = ctypes.create_string_buffer(128)
            if self.rank == 0:
                _lib.check(_lib.lib().cx_comm_unique_id(idbuf))
            ids = self.allgather_obj(bytes(idbuf.raw) if self.rank == 0 else None)
            _lib.check(_lib.lib().cx_comm_init(ctx.h, self.rank, self.size, ids[0]))
            # establish the NVLink peer connections now (NCCL connects lazily on first use)
            w = np.zeros(8, np.uint8)
            self.exchange_host({p: w for p in range(self.size) if p != self.rank},
                               {p: 8 for p in range(self.size) if p != self.rank})
            self._nccl_ready = os.environ.get("CX_ALLGATHER", "nccl") != "gloo"

    # -- small objects ------------------------------------------------------
    def allgather_obj(self, obj):
        if self.size == 1:
            return [obj]
        if self._nccl_ready:
            return self._allgather_obj_nccl(obj)
        out = [None] * self.size
        self._dist.all_gather_object(out, obj)
        return out

    def _allgather_obj_nccl(self, obj):
        """The pickled object of every rank through the library's NCCL communicator (two grouped exchanges: sizes,
        then the blobs): the tree layer gathers small metadata four or five times per distributed setInterpData,
        and a gloo all_gather_object costs several milliseconds each at 8 ranks (CX_ALLGATHER=gloo keeps it)."""
        import pickle
        blob = np.frombuffer(pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL), dtype=np.uint8)
        peers = [p for p in range(self.size) if p != self.rank]
        mysize = np.array([blob.size], np.int64).view(np.uint8)
        sizes = self._exchange_via_nccl({p: mysize for p in peers}, {p: 8 for p in peers}, np.uint8)
        nbytes = {p: int(sizes[p].view(np.int64)[0]) for p in peers}
        got = self._exchange_via_nccl({p: blob for p in peers}, nbytes, np.uint8)
        return [obj if p == self.rank else pickle.loads(got[p].tobytes()) for p in range(self.size)]

    def barrier(self):
        if self.size == 1:
            return
        if self.device:
            from . import _lib
            _lib.check(_lib.lib().cx_comm_barrier(self.ctx.h))
        else:
            self._dist.barrier()

    def allreduce_max(self, v):
        return max(self.allgather_obj(float(v)))

    def allreduce_sum(self, v):
        return sum(self.allgather_obj(v))

    # -- bulk data ------------------------------------------------------------
    def exchange_host(self, send, recv_sizes, dtype=np.uint8):
        """send: {peer: 1-D numpy array}; recv_sizes: {peer: n}.  Returns {peer: array}.
        Host path (gloo) — used when the communicator has no device, otherwise
        staged through device buffers and NCCL."""
        out = {}
        if self.size == 1:
            return out
        if self.device:
            return self._exchange_via_nccl(send, recv_sizes, dtype)
        import torch
        reqs = []
        keep = []
        for p, n in sorted(recv_sizes.items()):
            buf = torch.empty(int(n), dtype=torch.uint8 if np.dtype(dtype) == np.uint8 else torch.float64)
            out[p] = buf
            if n > 0:
                reqs.append(self._dist.irecv(buf, src=p))
        for p, a in sorted(send.items()):
            if a.size > 0:
                t = torch.from_numpy(np.ascontiguousarray(a))
                keep.append(t)
                reqs.append(self._dist.isend(t, dst=p))
        for r in reqs:
            r.wait()
        return {p: b.numpy() for p, b in out.items()}

    def _exchange_via_nccl(self, send, recv_sizes, dtype):
        from . import _lib
        import ctypes
        ctx = self.ctx
        isz = np.dtype(dtype).itemsize
        peers = sorted(set(send.keys()) | set(recv_sizes.keys()))
        sb, rb, sc, rc, hold = [], [], [], [], []
        for p in peers:
            a = send.get(p)
            ns = 0 if a is None else int(a.size) * isz
            nr = int(recv_sizes.get(p, 0)) * isz
            ds = ctx.alloc(max(ns, 8)); dr = ctx.alloc(max(nr, 8))
            hold.append((ds, dr))
            if ns > 0:
                ctx.upload(ds, np.ascontiguousarray(a))
            sb.append(ds); rb.append(dr); sc.append(ns); rc.append(nr)
        n = len(peers)
        _lib.check(_lib.lib().cx_exchange_bytes_dev(ctx.h, n, (ctypes.c_int * n)(*peers), _lib.ptr_array(sb),
                                                   (ctypes.c_longlong * n)(*sc), _lib.ptr_array(rb),
                                                   (ctypes.c_longlong * n)(*rc)))
        ctx.sync()
        out = {}
        for p, (ds, dr), nr in zip(peers, hold, rc):
            if p in recv_sizes:
                h = np.empty(nr // isz, dtype=dtype)
                if nr > 0:
                    ctx.download(h, dr)
                out[p] = h
            ctx.free(ds); ctx.free(dr)
        return out

    def exchange_dev(self, peers, send_ptrs, send_counts, recv_ptrs, recv_counts):
        """Grouped ncclSend/ncclRecv of doubles between device buffers (asynchronous)."""
        from . import _lib
        import ctypes
        n = len(peers)
        if n == 0:
            return
        _lib.check(_lib.lib().cx_exchange_dev(self.ctx.h, n, (ctypes.c_int * n)(*peers), _lib.ptr_array(send_ptrs),
                                              (ctypes.c_longlong * n)(*send_counts), _lib.ptr_array(recv_ptrs),
                                              (ctypes.c_longlong * n)(*recv_counts)))
