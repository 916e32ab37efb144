"""Four-step NTT over BN254 Fr sharded across the GPUs of one box: one process per GPU
(torch.distributed), ONE exchange step.

n = n1 * n2 (n1 = 2^floor(k/2) rows, n2 columns of the natural-order matrix x[i1*n2 + i2]), world size G:

  forward                                             layout on rank g
  ----------------------------------------------------------------------------------------------------------------
  in   A[i1][i2l]          i2 = g*c2 + i2l            column block: n1 x c2 (c2 = n2/G), row-major
  1    transpose                                      At[i2l][i1]
  2    c2 batched n1-point NTTs (omega^n2)            Yt[i2l][j1]
  3    exchange: * omega^(i2*j1), transpose,          Z[j1l][i2]   on rank h = j1 / r1  (r1 = n1/G, j1 = h*r1 + j1l)
       column block j1 in rank h's range -> rank h
  4    r1 batched n2-point NTTs (omega^n1)            X[j1l][j2] = DFT(x)[j1 + n1*j2]
  out  row block: r1 x n2, row-major; output index j lives on rank (j mod n1) / r1

The inverse runs the steps backwards (row INTTs, exchange with omega^-(i2*j1), column INTTs, transpose) and takes the
forward's output layout back to the forward's input layout, so forward -> pointwise work -> inverse needs no re-shuffle.

Step 3 has two transports:
  "p2p"   the exchange kernel stores over NVLink straight into the peers' Z buffers (torch symmetric memory gives the
          peer-mapped pointers): twiddle, transpose and all-to-all are ONE kernel, no staging copy, no unpack;
  "nccl"  the kernel writes per-destination staging chunks, dist.all_to_all_single moves them, a strided copy unpacks.
The local steps are the C-ABI calls panda_ntt_batch_execute_bn254_v1 / panda_ntt_exchange_bn254
(include/panda_interface.h); the CPU (gloo) tests inject host stand-ins for them.

The reference has no multi-GPU path and its single-GPU NTT is a stub (fft.cu:117-168 compiled out); this is new
functionality following the transform definition of panda_ntt_execute_bn254_v1.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_MONT = (1 << 256) % BN254_R
_MONT_INV = pow(_MONT, -1, BN254_R)


def fr_pow_mont(omega32, e: int) -> bytes:
    """omega^e for a Montgomery-form BN254 Fr element given as 32 little-endian bytes (host helper for the sub-roots)."""
    w = int.from_bytes(bytes(omega32), "little") * _MONT_INV % BN254_R
    return (pow(w, e, BN254_R) * _MONT % BN254_R).to_bytes(32, "little")


def split_log(log_n: int, world: int) -> tuple[int, int]:
    """(log n1, log n2) with n1 <= n2 and world | n1."""
    l1 = log_n // 2
    l2 = log_n - l1
    lg = world.bit_length() - 1
    if world & (world - 1) or l1 < lg:
        raise ValueError(f"world size {world} must be a power of two dividing n1 = 2^{l1}")
    return l1, l2


# ---- layout helpers (host side; tests, examples) ---------------------------------------------------------------------

def column_block(x_natural, log_n: int, rank: int, world: int):
    """rank's forward-input shard of a natural-order array of 2^log_n 32-byte elements (numpy uint8)."""
    l1, l2 = split_log(log_n, world)
    n1, n2 = 1 << l1, 1 << l2
    c2 = n2 // world
    return x_natural.reshape(n1, n2, 32)[:, rank * c2:(rank + 1) * c2].reshape(-1).copy()


def row_block_indices(log_n: int, rank: int, world: int):
    """natural-order output indices held by `rank` after forward(), in storage order: j = j1 + n1*j2."""
    import numpy as np

    l1, l2 = split_log(log_n, world)
    n1, n2 = 1 << l1, 1 << l2
    r1 = n1 // world
    j1 = np.arange(rank * r1, (rank + 1) * r1, dtype=np.int64)[:, None]
    j2 = np.arange(n2, dtype=np.int64)[None, :]
    return (j1 + n1 * j2).reshape(-1)


# ---- CUDA local steps -----------------------------------------------------------------------------------------------

class _CudaOps:
    def __init__(self):
        from . import gpu_ffi as ffi

        self.ffi = ffi

    def batch_ntt(self, d_src: int, d_dst: int, log_m: int, batch: int, omega32: bytes, inverse: bool, stream: int) -> int:
        """returns the pointer (d_src or d_dst) that holds the result"""
        ffi = self.ffi
        om = (C.c_uint8 * 32).from_buffer_copy(omega32)
        flag = C.c_uint(0)
        cfg = ffi.NttconfigurationV1(ffi.PandaMemPool(None), ffi.PandaStream(stream or None), d_src, d_dst, C.addressof(om), log_m, C.pointer(flag))
        rc = ffi.lib.panda_ntt_batch_execute_bn254_v1(cfg, batch, 1 if inverse else 0)
        if rc != 0:
            raise ffi.PandaGpuError("NTTExecuteErr", rc)
        return d_dst if flag.value else d_src

    def exchange(self, d_src: int, log_rows: int, log_cols: int, row_offset: int, omega32: Optional[bytes], log_n: int, inverse: bool,
                 dst_ptrs, ld: int, col_offset: int, stream: int) -> None:
        ffi = self.ffi
        om = (C.c_uint8 * 32).from_buffer_copy(omega32) if omega32 is not None else None
        arr = (C.c_void_p * len(dst_ptrs))(*dst_ptrs)
        cfg = ffi.NttExchangeConfiguration(ffi.PandaStream(stream or None), d_src, log_rows, log_cols, row_offset, log_n,
                                           C.addressof(om) if om is not None else None, 1 if inverse else 0, len(dst_ptrs), arr, ld, col_offset)
        rc = ffi.lib.panda_ntt_exchange_bn254(C.byref(cfg))
        if rc != 0:
            raise ffi.PandaGpuError("NTTExecuteErr", rc)


class ShardedNtt:
    """forward(x) / inverse(y) on this rank's shard (uint8 torch tensor of n/G * 32 bytes on the backend's device).

    transport: "p2p" (NVLink stores into peer buffers, needs torch symmetric memory), "nccl" (staging + all_to_all_single; also
    what the gloo tests use) or "auto" (p2p when the rendezvous succeeds).  ops: object with batch_ntt / exchange (defaults to
    the CUDA entry points).
    """

    def __init__(self, log_n: int, omega32, group=None, transport: str = "auto", ops=None, device=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.log_n = log_n
        self.l1, self.l2 = split_log(log_n, self.world)
        self.lg = self.world.bit_length() - 1
        self.omega = bytes(omega32)
        self.omega_rows = fr_pow_mont(self.omega, 1 << self.l2)      # order n1: the column transforms
        self.omega_cols = fr_pow_mont(self.omega, 1 << self.l1)      # order n2: the row transforms
        self.ops = ops or _CudaOps()
        self.local_elems = (1 << log_n) // self.world
        self.device = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if ops is None else torch.device("cpu"))
        nbytes = self.local_elems * 32
        self.buf_a = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.buf_b = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.symm = None
        self.transport = "nccl"
        if transport in ("p2p", "auto") and self.world > 1 and ops is None:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                self.recv = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
                self.symm = symm_mem.rendezvous(self.recv, group=group if group is not None else dist.group.WORLD)
                self.peer_ptrs = [int(p) for p in self.symm.buffer_ptrs]
                self.transport = "p2p"
            except Exception:
                if transport == "p2p":
                    raise
                self.symm = None
        if self.symm is None:
            self.recv = torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    # -- step 3 --------------------------------------------------------------------------------------------------------
    def _exchange(self, p, q, log_rows: int, log_cols: int, inverse: bool, stream: int):
        """p: tensor holding the (2^log_rows x 2^log_cols) local matrix whose global row offset is rank * rows; q: free scratch.
        Afterwards self.recv holds the (cols/G x rows*G) matrix: element (r, c) * omega^(+-(rank*rows + r) * c) of rank s sits at
        [c mod (cols/G)][s*rows + r] on rank c / (cols/G).  p and q are both free afterwards."""
        dist = self.dist
        G, g = self.world, self.rank
        rows, cols = 1 << log_rows, 1 << log_cols
        pc = cols // G
        if G == 1:
            self.ops.exchange(p.data_ptr(), log_rows, log_cols, 0, self.omega, self.log_n, inverse, [self.recv.data_ptr()], rows, 0, stream)
        elif self.transport == "p2p":
            self.symm.barrier(channel=0)             # every peer is done with its recv buffer from the previous use
            self.ops.exchange(p.data_ptr(), log_rows, log_cols, g * rows, self.omega, self.log_n, inverse, self.peer_ptrs, rows * G, g * rows, stream)
            self.symm.barrier(channel=1)             # all stores have landed
        else:
            chunk = pc * rows * 32
            self.ops.exchange(p.data_ptr(), log_rows, log_cols, g * rows, self.omega, self.log_n, inverse,
                              [q.data_ptr() + h * chunk for h in range(G)], rows, 0, stream)
            dist.all_to_all_single(p, q, group=self.group)      # p = [source rank][pc][rows]
            self.recv.view(pc, G, rows * 32).copy_(p.view(G, pc, rows * 32).permute(1, 0, 2))
        return self.recv

    def _stream(self, stream):
        """(raw handle, context manager): torch's own work of a call (copies, all_to_all, the symmetric-memory barriers) must be ordered on
        the same stream as the C-ABI kernels, so a foreign handle is entered as torch's current stream for the duration of the call"""
        import contextlib

        if self.device.type != "cuda":
            return int(stream or 0), contextlib.nullcontext()
        cur = self.torch.cuda.current_stream(self.device).cuda_stream
        if stream is None or int(stream) == cur:
            return cur, contextlib.nullcontext()
        return int(stream), self.torch.cuda.stream(self.torch.cuda.ExternalStream(int(stream), device=self.device))

    def _pick(self, ptr: int):
        for t in (self.buf_a, self.buf_b, self.recv):
            if t.data_ptr() == ptr:
                return t
        raise AssertionError("result pointer is none of the work buffers")

    def forward(self, x, stream=None, marks=None):
        """x: this rank's column block (module docstring); never written.  Returns this rank's row block in a buffer owned by this
        object (valid until the next forward / inverse call).  Work is queued on torch's current stream.
        marks: optional list; (phase name, CUDA event recorded on that stream at the END of the phase) pairs are appended to it, the first
        one ("start") before any work -- how bench.py splits the transform's time into its four steps (CUDA devices only)."""
        l1, l2, lg, ops = self.l1, self.l2, self.lg, self.ops
        stream, ctx = self._stream(stream)
        if x.numel() != self.local_elems * 32:
            raise ValueError("shard size mismatch")
        A, B = self.buf_a, self.buf_b

        def mark(name):
            if marks is not None and self.device.type == "cuda":
                ev = self.torch.cuda.Event(enable_timing=True)
                ev.record(self.torch.cuda.current_stream(self.device))
                marks.append((name, ev))

        with ctx:
            mark("start")
            ops.exchange(x.data_ptr(), l1, l2 - lg, 0, None, 0, False, [A.data_ptr()], 1 << l1, 0, stream)            # 1. At[i2l][i1]
            mark("transpose")
            p = self._pick(ops.batch_ntt(A.data_ptr(), B.data_ptr(), l1, 1 << (l2 - lg), self.omega_rows, False, stream))   # 2. Yt[i2l][j1]
            mark("column_transforms")
            q = B if p is A else A
            z = self._exchange(p, q, l2 - lg, l1, False, stream)                                                       # 3. Z[j1l][i2]
            mark("exchange")
            out = self._pick(ops.batch_ntt(z.data_ptr(), A.data_ptr(), l2, 1 << (l1 - lg), self.omega_cols, False, stream))  # 4. X[j1l][j2]
            mark("row_transforms")
            return out

    def inverse(self, y, stream=None):
        """y: this rank's row block (the layout forward() returns).  Returns the column block of (1/n) DFT_{omega^-1}."""
        l1, l2, lg, ops = self.l1, self.l2, self.lg, self.ops
        stream, ctx = self._stream(stream)
        if y.numel() != self.local_elems * 32:
            raise ValueError("shard size mismatch")
        A, B = self.buf_a, self.buf_b
        with ctx:
            if y.data_ptr() == B.data_ptr():
                src, scratch = B, A
            else:
                if y.data_ptr() != A.data_ptr():
                    A.copy_(y)                       # the transforms ping-pong between their two buffers: keep the caller's (or recv) intact
                src, scratch = A, B
            p = self._pick(ops.batch_ntt(src.data_ptr(), scratch.data_ptr(), l2, 1 << (l1 - lg), self.omega_cols, True, stream))   # Z[j1l][i2]
            q = B if p is A else A
            w = self._exchange(p, q, l1 - lg, l2, True, stream)                                                        # W[i2l][j1]
            p = self._pick(ops.batch_ntt(w.data_ptr(), A.data_ptr(), l1, 1 << (l2 - lg), self.omega_rows, True, stream))    # At[i2l][i1]
            ops.exchange(p.data_ptr(), l2 - lg, l1, 0, None, 0, False, [B.data_ptr()], 1 << (l2 - lg), 0, stream)       # A[i1][i2l]
            return B
