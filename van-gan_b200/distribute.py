"""Data-parallel plumbing: the role tf.distribute.MirroredStrategy plays in the reference
(main.py:22, vangan.py:86,459-490), as one process per GPU over torch.distributed.

The train step shards by batch only (InstanceNorm is per sample), so the one exchange step is the
SUM all-reduce of each network's flat gradient buffer — issued right after that network's backward
sweep so it overlaps the next network's backward — plus one 10-float all-reduce of the result dict.
"""
import contextlib
import os

import torch
import torch.distributed as dist


BUCKET_ELEMS = 6 * 1024 * 1024 + 256 * 1024      # ~25 MB fp32 messages


class _Done:
    """Handle of an all-reduce enqueued on the communication stream: wait() orders the CURRENT stream after it (no host sync)."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class Strategy:
    """Minimal MirroredStrategy look-alike: `.num_replicas_in_sync`, `.scope()`, `.run()`, `.reduce()`.

    CUDA tensors are exchanged through the library's own communicator (`vg_comm_*`, NCCL resolved at run time) on a dedicated
    communication stream ordered by events -- the same calls are captured into the CUDA graph of the train step.  torch.distributed
    supplies the rendezvous (rank / world size, the broadcast of the 128-byte NCCL id) and carries CPU tensors (gloo tests)."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.num_replicas_in_sync = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.comm = None            # vg_comm* (created on the first CUDA exchange)
        self.comm_stream = None
        self.use_vg_comm = os.environ.get("VG_COMM", "1") != "0"

    # ------------------------------------------------------------------ the library's communicator
    def _ensure_comm(self):
        if self.comm is not None:
            return True
        if not (self.use_vg_comm and self.enabled and self.num_replicas_in_sync > 1 and torch.cuda.is_available()):
            return False
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        if not L.vg_comm_available():
            raise _lib.VgError("vg_comm: libnccl.so.2 could not be resolved (set VG_NCCL_LIB)")
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            rc = L.vg_comm_unique_id(buf)
            if rc != 0:
                raise _lib.VgError("vg_comm_unique_id failed (%d)" % rc)
        box = [buf.raw]
        dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        handle = C.c_void_p()
        rc = L.vg_comm_init(C.byref(handle), box[0], self.num_replicas_in_sync, self.rank, torch.cuda.current_device())
        if rc != 0:
            raise _lib.VgError("vg_comm_init failed (%d)" % rc)
        self.comm, self.comm_stream, self._L = handle, torch.cuda.Stream(), L
        return True

    def destroy(self):
        if self.comm is not None:
            torch.cuda.synchronize()
            self._L.vg_comm_destroy(self.comm)
            self.comm = None

    @contextlib.contextmanager
    def scope(self):
        yield self

    def run(self, fn, args=(), kwargs=None):
        return fn(*args, **(kwargs or {}))

    def all_reduce_async(self, tensor, bucket_elems=BUCKET_ELEMS):
        """SUM all-reduce of a flat fp32 buffer, in ~25 MB messages; returns a handle with .wait() (None when single replica).
        CUDA: enqueued on the communication stream after everything the current stream has enqueued so far, so it overlaps
        whatever the current stream enqueues next (the next network's backward sweep)."""
        if not self.enabled or self.num_replicas_in_sync == 1:
            return None
        if tensor.is_cuda and self._ensure_comm():
            assert tensor.dtype == torch.float32 and tensor.is_contiguous()
            ready = torch.cuda.Event()
            ready.record()
            self.comm_stream.wait_event(ready)
            rc = self._L.vg_comm_allreduce_bucket(self.comm, tensor.data_ptr(), tensor.numel(), int(bucket_elems), self.comm_stream.cuda_stream)
            if rc != 0:
                from . import _lib
                raise _lib.VgError("vg_comm_allreduce_bucket failed (%d)" % rc)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
            return _Done(done)
        return dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def reduce(self, op, value, axis=None):
        """strategy.reduce(ReduceOp.SUM, v, axis=None) on a tensor (fp64 CUDA tensors: vg_comm_reduce_scalars on the current stream)."""
        if self.enabled and self.num_replicas_in_sync > 1:
            if value.is_cuda and value.dtype == torch.float64 and value.is_contiguous() and self._ensure_comm():
                rc = self._L.vg_comm_reduce_scalars(self.comm, value.data_ptr(), value.numel(), torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    from . import _lib
                    raise _lib.VgError("vg_comm_reduce_scalars failed (%d)" % rc)
            elif value.is_cuda and self._ensure_comm():
                h = self.all_reduce_async(value.view(-1), bucket_elems=0)
                h.wait()
            else:
                dist.all_reduce(value, op=dist.ReduceOp.SUM, group=self.group)
        return value


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local
