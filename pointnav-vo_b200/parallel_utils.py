"""Host-side data-parallel helpers (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in
the CPU tests).  The VO path has exactly two exchanges per training step:
  1. the flat fp32 gradient bucket -- ONE all-reduce (SUM; the 1/world factor is applied by the loss-gradient
     kernel, so the result is the gradient of the mean loss over the concatenated batch);
  2. the packed RunningMeanAndVar statistics (sum, sum of squares per channel, fp64) -- ONE all-reduce instead of
     the reference's three (model_utils/running_mean_and_var.py:28-38)."""
import torch
import torch.distributed as dist


def allreduce_flat_bucket(flat, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
    return flat


def merge_input_stats(packed, n_local, pix, group=None):
    """packed: [2C] (sum, sumsq interleaved) of this rank's batch.  Returns the (mean, population variance) of the
    batch concatenated over all ranks -- what RunningMeanAndVar computes with its all-reduces."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    allreduce_flat_bucket(packed, group)
    n = float(n_local * world * pix)
    s = packed.view(-1, 2)
    mean = s[:, 0] / n
    var = (s[:, 1] / n - mean * mean).clamp_min(0)
    return mean, var


def peer_slice(n_pad, rank, world):
    """[lo, hi) elements of a bucket of n_pad floats (multiple of 4) owned by `rank`: the slice pnvo_peer_reduce_adam reduces and
    updates there (csrc/peer_reduce.cu: slice4 = ceil(n4 / world) float4 per rank, the last ranks may own less or nothing)."""
    n4 = n_pad // 4
    s4 = (n4 + world - 1) // world
    return min(4 * s4 * rank, n_pad), min(4 * s4 * (rank + 1), n_pad)


class _RawCuda:
    """A caller-owned device range exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerRegion:
    """One zero-filled, peer-mappable device region per rank (libpnvo's pnvo_peer_alloc: cudaMalloc + CUDA IPC handle) and
    the mappings of every other rank's region: NVLink loads / stores between the processes of one node.  Construction is
    COLLECTIVE over `group`; a failure on any rank raises PnvoError on every rank."""

    def __init__(self, nbytes, device, group=None):
        import ctypes
        import socket

        from . import lib as L

        self.L = L
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if not 2 <= self.world <= 8:
            raise L.PnvoError("PeerRegion: 2..8 ranks on one node")
        self.dev = torch.device(device)
        self.nbytes = int(nbytes)
        lib = L.load()
        base = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        ok = 1
        try:  # failures are agreed on collectively below: no rank may leave before the all_gather
            with torch.cuda.device(self.dev):
                L.check(lib.pnvo_peer_alloc(self.nbytes, ctypes.byref(base), handle))
        except Exception as e:  # noqa: BLE001
            ok = 0
            self._error = e
        self._base = base.value
        # the host name rides along: IPC handles only mean something on the same node
        host = socket.gethostname().encode()[:63].ljust(64, b"\0")
        mine = torch.tensor(list(bytes(handle)) + list(host), dtype=torch.uint8, device=self.dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        self.bases = [None] * self.world
        self._opened = []
        try:
            for p, g in enumerate(gathered):
                if not ok:
                    break
                raw = bytes(g.cpu().tolist())
                if raw[64:] != host:
                    raise L.PnvoError("PeerRegion: ranks are not on one node")
                if p == self.rank:
                    self.bases[p] = self._base
                    continue
                h = (ctypes.c_ubyte * 64)(*raw[:64])
                out = ctypes.c_void_p()
                with torch.cuda.device(self.dev):
                    L.check(lib.pnvo_peer_open(h, ctypes.byref(out)))
                self.bases[p] = out.value
                self._opened.append(out.value)
        except Exception as e:  # noqa: BLE001 -- the outcome is agreed on collectively below
            ok = 0
            self._error = e
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close()
            raise L.PnvoError(f"PeerRegion: peer mapping failed on at least one rank ({getattr(self, '_error', 'another rank')})")
        dist.barrier(group=group)

    def view(self, byte_offset, n, typestr):
        return torch.as_tensor(_RawCuda(self._base + byte_offset, n, typestr), device=self.dev)

    def pointer_array(self, byte_offset):
        import ctypes

        return (ctypes.c_void_p * self.world)(*[b + byte_offset for b in self.bases])

    def close(self):
        lib = self.L.load()
        for p in getattr(self, "_opened", []):
            lib.pnvo_peer_close(p)
        self._opened = []
        if getattr(self, "_base", None):
            lib.pnvo_peer_free(self._base)
            self._base = None


class PeerBuckets(PeerRegion):
    """This rank's flat parameter and gradient buckets inside ONE peer-mappable region.  `reduce_adam()` launches
    libpnvo's fused reduce-scatter + Adam + all-gather kernel (csrc/peer_reduce.cu) instead of `all_reduce(grads)`
    followed by an optimiser kernel: the replacement, for the flat buckets of the VO trainer, of
    DistributedDataParallel's reducer + torch.optim.Adam.step() (rl/ddppo/algo/ddppo.py:55-96).

    Region layout (fp32 elements): [params n_pad][grads n_pad][flags 64 x u32]."""

    FLAG_WORDS = 64

    def __init__(self, n, device, group=None):
        self.n = int(n)
        self.n_pad = (self.n + 3) // 4 * 4
        super().__init__((2 * self.n_pad + self.FLAG_WORDS) * 4, device, group)
        self.params = self.view(0, self.n_pad, "<f4")
        self.grads = self.view(4 * self.n_pad, self.n_pad, "<f4")
        self.flags = self.view(8 * self.n_pad, self.FLAG_WORDS, "<i4")
        self._params_arr = self.pointer_array(0)
        self._grads_arr = self.pointer_array(4 * self.n_pad)
        self._flags_arr = self.pointer_array(8 * self.n_pad)
        self.seq = 0

    def reduce_adam(self, m, v, step, lr, beta1, beta2, eps):
        """m, v: this rank's fp32 moment buffers (n_pad elements; only this rank's slice is touched)."""
        L = self.L
        assert m.numel() >= self.n_pad and v.numel() >= self.n_pad
        self.seq += 1
        L.check(L.load().pnvo_peer_reduce_adam(self._grads_arr, self._params_arr, self._flags_arr, L.ptr(m), L.ptr(v),
                                               self.n_pad, self.rank, self.world, self.seq, lr, beta1, beta2, eps,
                                               int(step), L.stream_ptr(self.dev)))

    def slice_range(self):
        """[lo, hi) element range of the bucket whose Adam moments live on this rank."""
        return peer_slice(self.n_pad, self.rank, self.world)

    def timed_out(self):
        return bool(self.flags[17].item() != 0)


class PeerSmallSum(PeerRegion):
    """All-reduce (sum, fixed rank order) of small fp64 vectors through peer memory: `sum_(t)` replaces
    `dist.all_reduce(t)` for the packed RunningMeanAndVar statistics inside the training step.  One tiny CTA that
    co-resides with the persistent convolution kernels (an NCCL kernel spinning for the slowest rank does not: it
    blocks the placement of their 148th CTA -- measured 0.21 ms per step at 2 GPUs).

    Region layout: [slots world x 2 x 1024 f64][flags 64 x u32]."""

    MAX_N = 1024

    def __init__(self, device, group=None):
        world = dist.get_world_size(group)
        self._slot_bytes = world * 2 * self.MAX_N * 8
        super().__init__(self._slot_bytes + 256, device, group)
        self.flags = self.view(self._slot_bytes, 64, "<i4")
        self._slots_arr = self.pointer_array(0)
        self._flags_arr = self.pointer_array(self._slot_bytes)
        self.seq = 0

    def sum_(self, t):
        L = self.L
        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous() or t.numel() > self.MAX_N:
            raise L.PnvoError("PeerSmallSum.sum_: contiguous fp64 CUDA tensor of at most 1024 elements")
        self.seq += 1
        L.check(L.load().pnvo_peer_sum_f64(self._slots_arr, self._flags_arr, L.ptr(t), t.numel(), self.rank, self.world,
                                           self.seq, L.stream_ptr(self.dev)))
        return t

    def timed_out(self):
        return bool(self.flags[17].item() != 0)
