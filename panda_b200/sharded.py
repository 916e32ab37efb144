"""MSM sharded by point range over the GPUs of one box: one process per GPU (torch.distributed), no data-path
collective -- each rank runs the full single-GPU pipeline on its contiguous slice of the points and only the 96-byte
(144-byte for BLS12-377) Jacobian partials are all-gathered and summed.

The reference has no multi-GPU path (device 0 is hard-coded, msm_cuda.cuh:554-555); this is new functionality behind
the additions of include/panda_interface.h (panda_msm_execute_*_n, panda_msm_combine_*).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced point range [lo, hi) owned by `rank`; ranges tile [0, n) exactly."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def result_bytes(curve: int) -> int:
    return 144 if curve == 1 else 96


def _cuda_local_msm(curve: int):
    import ctypes as C
    from . import gpu_ffi as ffi

    def run(d_bases: int, d_scalars: int, n_local: int, d_out: int, stream: int, pool: int) -> None:
        cfg = ffi.MSMConfiguration(ffi.PandaMemPool(pool or None), ffi.PandaStream(stream or None), d_bases, d_scalars, d_out, 0,
                                   ffi.PandaMSMResultCoordinateType.Jacobian)
        fn = ffi.lib.panda_msm_execute_bls12_377_n if curve == 1 else ffi.lib.panda_msm_execute_bn254_n
        rc = fn(cfg, C.c_size_t(n_local))
        if rc != 0:
            raise ffi.PandaGpuError("SchedulingErr", rc)

    return run


def _cuda_combine(curve: int):
    from . import gpu_ffi as ffi

    def run(d_partials: int, count: int, d_out: int, coord: int, stream: int) -> None:
        fn = ffi.lib.panda_msm_combine_bls12_377 if curve == 1 else ffi.lib.panda_msm_combine_bn254
        rc = fn(d_partials, count, d_out, coord, ffi.PandaStream(stream or None))
        if rc != 0:
            raise ffi.PandaGpuError("SchedulingErr", rc)

    return run


class ShardedMsm:
    """result = sum over ranks of MSM(local points); every rank ends up with the combined result.

    local_msm(d_bases, d_scalars, n_local, d_out, stream, pool) and combine(d_partials, count, d_out, coord, stream) default
    to the CUDA entry points; the CPU (gloo) tests inject host stand-ins to exercise the partition / gather logic.
    """

    def __init__(self, curve: int = 0, group=None, local_msm: Optional[Callable] = None, combine: Optional[Callable] = None):
        import torch.distributed as dist

        self.dist = dist
        self.curve = curve
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.local_msm = local_msm or _cuda_local_msm(curve)
        self.combine = combine or _cuda_combine(curve)

    def run(self, bases_t, scalars_t, n_local: int, coord: int = 0, stream=None, pool: int = 0):
        """bases_t / scalars_t: uint8 torch tensors holding this rank's slice (on the device the backend works on).
        Returns a uint8 tensor with the 3-coordinate result, identical on every rank.
        stream: raw CUDA stream handle or None (= torch's current stream).  The kernels AND torch's own work (allocations,
        the all-gather) are ordered on that one stream: a foreign handle is entered as torch's current stream for the call."""
        import contextlib

        import torch

        dev = bases_t.device
        ctx = contextlib.nullcontext()
        if dev.type == "cuda":
            cur = torch.cuda.current_stream(dev).cuda_stream
            if stream is None:
                stream = cur
            elif int(stream) != cur:
                ctx = torch.cuda.stream(torch.cuda.ExternalStream(int(stream), device=dev))
        stream = int(stream or 0)
        nb = result_bytes(self.curve)
        with ctx:
            partial = torch.empty(nb, dtype=torch.uint8, device=dev)
            self.local_msm(bases_t.data_ptr(), scalars_t.data_ptr(), n_local, partial.data_ptr(), stream, pool)
            if self.world > 1:
                gathered = torch.empty(self.world * nb, dtype=torch.uint8, device=dev)
                self.dist.all_gather_into_tensor(gathered, partial, group=self.group)   # the only exchange: world x 96 bytes
            else:
                gathered = partial
            out = torch.empty(nb, dtype=torch.uint8, device=dev)
            self.combine(gathered.data_ptr(), self.world, out.data_ptr(), coord, stream)
        return out


def _cuda_class_msm(curve: int):
    import ctypes as C
    from . import gpu_ffi as ffi

    def run(d_bases: int, d_scalars: int, n: int, class_count: int, class_index: int, d_out: int, stream: int, pool: int) -> None:
        cfg = ffi.MSMConfiguration(ffi.PandaMemPool(pool or None), ffi.PandaStream(stream or None), d_bases, d_scalars, d_out, 0,
                                   ffi.PandaMSMResultCoordinateType.Jacobian)
        fn = ffi.lib.panda_msm_execute_bls12_377_class if curve == 1 else ffi.lib.panda_msm_execute_bn254_class
        rc = fn(cfg, C.c_size_t(n), class_count, class_index)
        if rc != 0:
            raise ffi.PandaGpuError("SchedulingErr", rc)

    return run


class ClassShardedMsm(ShardedMsm):
    """MSM sharded by BUCKET CLASS: every rank holds all n points and scalars (and the whole table of the cached bases) and adds up the digits
    whose bucket index is congruent to its rank modulo the world size (panda_msm_execute_*_class); the 96-byte partials are all-gathered and
    summed like ShardedMsm's.  For device-resident scalars on a box whose GPUs each have room for the whole job: the window width stays the
    single-GPU one and the bucket reduction is split, not repeated.  The world size must be a power of two."""

    def __init__(self, curve: int = 0, group=None, class_msm: Optional[Callable] = None, combine: Optional[Callable] = None):
        super().__init__(curve, group, local_msm=lambda *a: None, combine=combine)
        if self.world & (self.world - 1):
            raise ValueError("bucket-class shards need a power-of-two world size")
        self.class_msm = class_msm or _cuda_class_msm(curve)
        self.local_msm = lambda d_b, d_s, n, d_out, stream, pool: self.class_msm(d_b, d_s, n, self.world, self.rank, d_out, stream, pool)
