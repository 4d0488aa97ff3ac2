"""Slab-decomposed N-D FFT for single large grids sharded over the GPUs of one node.

Real-space arrays are sharded on axis 0 (``(n0/G, n1, ...)`` per rank); spectral arrays are kept in
the TRANSPOSED layout, sharded on axis 1 (``(n0, n1/G, ...)``), so that every N-D transform costs
exactly one all-to-all (NCCL over NVLink; gloo in the CPU tests).  ``lin_op``, ``u`` and every
engine buffer use the spectral layout, which makes the diagonal stepping kernels (K1/K2/K3) purely
local; only the three error-norm scalars and the transposes cross ranks.

On CUDA with power-of-two axis lengths the local transforms are the engine's own kernels
(``fused_nl``: strided-axis transforms of csrc/fft_axis.cuh around the fused last-axis kernel, physical
data left digit-reversed along the strided axes); other sizes and the CPU/gloo tests use ``torch.fft``.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class SlabFFT:
    """c2c FFT of a global ``shape`` grid whose first axis (real space) / second axis (spectral
    space) is split over the ranks of ``group``."""

    def __init__(self, shape: Sequence[int], group=None) -> None:
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) < 2:
            raise ValueError("slab decomposition needs at least a 2-D grid")
        n0, n1 = self.shape[:2]
        if n0 % self.world or n1 % self.world:
            raise ValueError(f"the first two grid dimensions {n0}, {n1} must be divisible by the world size {self.world}")
        self.rest = self.shape[2:]
        self.real_shape = (n0 // self.world, n1) + self.rest          # this rank's real-space slab
        self.spec_shape = (n0, n1 // self.world) + self.rest          # this rank's spectral pencil block
        self._peer = None                                             # symmetric buffers of the peer-memory exchange

    # -- layout helpers ---------------------------------------------------------------------
    def real_slice(self, a: torch.Tensor) -> torch.Tensor:
        """This rank's slab of a global real-space array."""
        m = self.shape[0] // self.world
        return a[self.rank * m:(self.rank + 1) * m].contiguous()

    def spec_slice(self, a: torch.Tensor) -> torch.Tensor:
        """This rank's block of a global spectral array (axis 1 sharded)."""
        m = self.shape[1] // self.world
        return a[:, self.rank * m:(self.rank + 1) * m].contiguous()

    def _exchange(self, a: torch.Tensor) -> torch.Tensor:
        """All-to-all of the leading-axis blocks of ``a`` (block g goes to rank g)."""
        if self.world == 1:
            return a
        out = torch.empty_like(a)
        if self._nccl() and a.shape[0] == self.world:
            # this rank's own block never leaves the GPU: a plain device copy instead of a self-send through the
            # communicator's channels (the block is 1/G of the data: half of it on two GPUs)
            for w in self._exchange_lists([out[g] for g in range(self.world)], [a[g] for g in range(self.world)]):
                w.wait()
            return out
        dist.all_to_all_single(out, a, group=self.group)
        return out

    def _nccl(self) -> bool:
        return dist.get_backend(self.group) == "nccl"

    def _exchange_lists(self, outs, ins):
        """Asynchronous all-to-all of per-peer CONTIGUOUS blocks (``ins[g]`` goes to rank g, ``outs[g]`` arrives from
        rank g); returns a list of work handles.  NCCL: one grouped send/recv launch on the communicator's stream, so
        the exchange of one chunk runs beside the transforms of another; other backends (gloo in the CPU tests):
        point-to-point operations."""
        if self._nccl():
            me = self.rank
            outs[me].copy_(ins[me])                              # own block: device copy on the stepping stream
            nothing = ins[me].new_empty(0)
            ins = [nothing if g == me else t for g, t in enumerate(ins)]
            outs = [nothing if g == me else t for g, t in enumerate(outs)]
            return [dist.all_to_all(outs, ins, group=self.group, async_op=True)]
        outs[self.rank].copy_(ins[self.rank])
        ops = []
        for g in range(self.world):
            if g != self.rank:
                peer = dist.get_global_rank(self.group, g) if self.group is not None else g
                ops.append(dist.P2POp(dist.isend, ins[g], peer, group=self.group))
                ops.append(dist.P2POp(dist.irecv, outs[g], peer, group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    # -- transforms -------------------------------------------------------------------------
    def forward(self, f: torch.Tensor, skip_last: bool = False) -> torch.Tensor:
        """real-space slab ``(n0/G, n1, ...)`` -> spectral block ``(n0, n1/G, ...)`` (unnormalised).
        ``skip_last``: the last axis is already in spectral space (fused row kernel did it)."""
        G, (n0, n1) = self.world, self.shape[:2]
        nd = len(self.shape)
        a = torch.fft.fftn(f, dim=tuple(range(1, nd - 1 if skip_last and nd > 2 else nd)))
        # chunk j of axis 1 goes to rank j
        a = a.reshape((n0 // G, G, n1 // G) + self.rest).permute((1, 0, 2) + tuple(range(3, nd + 1))).contiguous()
        b = self._exchange(a).reshape(self.spec_shape)        # blocks arrive ordered by source rank = axis-0 order
        return torch.fft.fft(b, dim=0)

    def inverse(self, s: torch.Tensor, skip_last: bool = False) -> torch.Tensor:
        """spectral block -> real-space slab (normalised like numpy.fft.ifftn).
        ``skip_last``: leave the last axis in spectral space (the fused row kernel transforms it)."""
        G, (n0, n1) = self.world, self.shape[:2]
        nd = len(self.shape)
        b = torch.fft.ifft(s, dim=0).reshape((G, n0 // G, n1 // G) + self.rest).contiguous()
        a = self._exchange(b)                                  # a[j] = my axis-0 planes of rank j's axis-1 chunk
        a = a.permute((1, 0, 2) + tuple(range(3, nd + 1))).reshape(self.real_shape)
        return torch.fft.ifftn(a, dim=tuple(range(1, nd - 1 if skip_last and nd > 2 else nd)))


    # -- exchange through peer memory ----------------------------------------------------------
    def _peer_setup(self, like: torch.Tensor) -> bool:
        """Symmetric buffers for the two exchanges of a 3-D evaluation (torch symmetric memory: every rank maps every
        rank's buffer over NVLink) and the address tables the scattering transforms take.  False -- and the NCCL
        route stays in use -- when RKS_SLAB_P2P=0, off CUDA/NCCL, for other ranks than 3-D power-of-two worlds, or
        when the rendezvous fails on this system."""
        if self._peer is not None:
            return self._peer is not False
        import os
        self._peer = False
        G, nd = self.world, len(self.shape)
        ok = (G > 1 and nd == 3 and like.is_cuda and self._nccl() and G & (G - 1) == 0
              and os.environ.get("RKS_SLAB_P2P", "1")[:1] != "0")
        try:
            if ok:
                import torch.distributed._symmetric_memory as symm_mem
                n0, n1, n2 = self.shape
                m, q = n0 // G, n1 // G
                numel = G * m * q * n2                                   # = the local block of the grid
                bufs = [symm_mem.empty(numel, dtype=torch.complex128, device=like.device) for _ in range(2)]
                group = self.group if self.group is not None else dist.group.WORLD
                hdls = [symm_mem.rendezvous(b, group) for b in bufs]
                block = m * q * n2 * 16                                  # bytes of one (source, destination) block
                tabs = [torch.tensor([int(h.buffer_ptrs[g]) + self.rank * block for g in range(G)], dtype=torch.int64,
                                     device=like.device) for h in hdls]
                # flags of the engine's own barrier kernel (rks_peer_barrier): G uint64 per rank, zero before first use
                flags = symm_mem.empty(32, dtype=torch.int64, device=like.device)
                flags.zero_()
                fh = symm_mem.rendezvous(flags, group)
                ftab = torch.tensor([int(fh.buffer_ptrs[g]) for g in range(G)], dtype=torch.int64, device=like.device)
                torch.cuda.synchronize(like.device)
                fh.barrier()                                             # every rank's flags are zero before anyone signals
                self._peer = {"bufs": bufs, "hdls": hdls, "tabs": tabs, "flags": flags, "fh": fh, "ftab": ftab, "epoch": 0,
                              "own_barrier": os.environ.get("RKS_PEER_BARRIER", "1")[:1] != "0"}
        except Exception as exc:                                          # noqa: BLE001
            import warnings
            warnings.warn(f"slab exchange through peer memory unavailable ({exc!r}): using NCCL all-to-all")
            ok = False
        # every rank must take the same route
        flag = torch.tensor([1 if self._peer else 0], dtype=torch.int32, device=like.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self._peer = False
        return self._peer is not False

    def _peer_barrier(self, which: int) -> None:
        """Order point between the ranks: the engine's one-block flag kernel (one NVLink round trip) or, with
        RKS_PEER_BARRIER=0, the symmetric-memory handle's own barrier."""
        pr = self._peer
        if not pr["own_barrier"]:
            pr["hdls"][which].barrier()
            return
        from ctypes import c_void_p
        from . import _abi
        pr["epoch"] += 1
        dev = pr["flags"].device
        _abi.check(_abi.lib.rks_peer_barrier(c_void_p(pr["ftab"].data_ptr()), self.world, self.rank, pr["epoch"],
                                             c_void_p(torch.cuda.current_stream(dev).cuda_stream)))

    def _fused_nl_peer(self, s: torch.Tensor, rows, axes, out: Optional[torch.Tensor]) -> torch.Tensor:
        """3-D ``fused_nl`` without a collective: the inverse transform over axis 0 stores its output rows straight
        into the x-plane owners' buffers, and the forward transform over axis 1 stores its rows straight into the
        y-chunk owners' buffers (``AxisFFT.scatter_``: the last level's stores go over NVLink), so each exchange is
        fused into the kernel that produces its data -- no extra pass over HBM, no NCCL kernel.  One barrier after
        each scattering kernel orders the ranks (a buffer is rewritten only after the barrier that follows its last
        read).  Same kernels and values as the NCCL route: bit-identical results."""
        G, (n0, n1, n2) = self.world, self.shape
        m, q = n0 // G, n1 // G
        bufs, hdls, tabs = self._peer["bufs"], self._peer["hdls"], self._peer["tabs"]
        # [inverse over x | exchange]: row p of the x axis belongs to rank p // m; lands as a[my rank][p % m][ky][z]
        axes[0].scatter_(s, tabs[0], 1, q * n2, 1, True)
        self._peer_barrier(0)
        a = bufs[0].view(G, m, q, n2)                                     # [source rank = ky chunk][my x planes][ky][z]
        axes[1].chunked_(a, True)
        rows(a, out=a)
        # [forward over y | exchange]: row p of the ky axis belongs to rank p // q; lands as b[my rank][x][p % q][z]
        axes[1].scatter_(a, tabs[1], m, n2, G, False)
        self._peer_barrier(1)
        return axes[0].forward_(bufs[1].view(self.spec_shape), 0, out=out)

    def fused_nl(self, s: torch.Tensor, rows, axes, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``F{ N( F^-1{ s } ) }`` with the engine's kernels only: ``axes[d]`` is the ``AxisFFT`` of grid axis d
        (d < nd-1), ``rows`` the fused last-axis kernel (inverse, pointwise N, forward).  Two exchanges: through peer
        memory, fused into the transforms (3-D grids on NVLink-connected GPUs), else two all-to-alls."""
        G, (n0, n1), nd = self.world, self.shape[:2], len(self.shape)
        if G > 1 and nd == 3 and self._peer_setup(s):
            return self._fused_nl_peer(s, rows, axes, out)
        chunks = self._pipeline_chunks()
        if chunks > 1:
            return self._fused_nl_pipelined(s, rows, axes, out, chunks)
        work = torch.empty_like(s)
        axes[0].inverse_(s, 0, out=work)                                   # s itself must stay intact
        a = self._exchange(work.reshape((G, n0 // G, n1 // G) + self.rest))
        if nd > 2:
            # a = [source rank g][my x planes][g's chunk of axis 1][rest]: axis 1 is transformed in this
            # chunk-major layout, so neither side of the exchange needs a transposing copy
            if nd == 3:
                a = a.reshape((G, n0 // G, n1 // G, self.rest[0]))
            axes[1].chunked_(a, True)
            for d in range(2, nd - 1):
                axes[d].inverse_(a, d + 1)
            rows(a, out=a)
            for d in range(nd - 2, 1, -1):
                axes[d].forward_(a, d + 1)
            axes[1].chunked_(a, False)
        else:
            # 2-D: axis 1 is the last (contiguous) axis and must be whole for the row kernel
            a = a.permute((1, 0, 2)).reshape(self.real_shape)              # a copy only when G > 1
            rows(a, out=a)
            a = a.reshape((n0 // G, G, n1 // G)).permute((1, 0, 2)).contiguous()
        b = self._exchange(a).reshape(self.spec_shape)
        return axes[0].forward_(b, 0, out=out)


    # -- overlapped exchange ------------------------------------------------------------------
    #: x-plane groups the exchange of a 3-D (or higher) evaluation is split into (RKS_SLAB_CHUNKS overrides; 1 = off).
    #: Measured on 2 B200s, 512^3 (profiles/r02f_cfg5_chunks*.json, r02i_cfg5_*.json): 35.8 ms per trial unchunked,
    #: 36.0-38.7 ms with 2-8 groups, whatever the SM margin left to NCCL or its stream priority -- the NCCL
    #: send/recv kernels and the transforms do not overlap in practice, so the pipeline is off by default.
    PIPELINE_CHUNKS = 1

    def _pipeline_chunks(self) -> int:
        """How many x-plane groups to pipeline: needs >= 3 dimensions, more than one rank and a divisor of the local
        plane count; the largest divisor <= the requested number is used."""
        import os
        want = int(os.environ.get("RKS_SLAB_CHUNKS", self.PIPELINE_CHUNKS))
        m = self.shape[0] // self.world
        if self.world == 1 or len(self.shape) < 3 or want <= 1:
            return 1
        c = min(want, m)
        while m % c:
            c -= 1
        return c

    def _fused_nl_pipelined(self, s: torch.Tensor, rows, axes, out: Optional[torch.Tensor], C: int) -> torch.Tensor:
        """``fused_nl`` with the two all-to-alls split into C groups of this rank's x planes ("chunk along
        the local axis to overlap with FFT passes").  After the axis-0 inverse, group c of every destination's planes
        is one contiguous block, so group c's exchange (NCCL stream) runs while group c-1 goes through the axis-1
        transform, the fused last-axis kernel and the axis-1 forward transform (stepping stream), and group c-2
        travels back.  Only the two axis-0 passes are outside the overlap.  Same kernels on the same values as the
        unchunked route: results are bit-identical."""
        G, (n0, n1), nd = self.world, self.shape[:2], len(self.shape)
        m = n0 // G
        mc = m // C
        work = torch.empty_like(s)
        axes[0].inverse_(s, 0, out=work)                                   # s itself must stay intact
        wv = work.reshape((G, m, n1 // G) + self.rest)                     # [destination g][its x plane][my y chunk]
        a = torch.empty((C, G, mc, n1 // G) + self.rest, dtype=s.dtype, device=s.device)
        b = torch.empty_like(s)
        bv = b.reshape((G, m, n1 // G) + self.rest)                        # [source g = x block][x plane][my y chunk]
        arriving = [self._exchange_lists([a[c, g] for g in range(G)], [wv[g, c * mc:(c + 1) * mc] for g in range(G)])
                    for c in range(C)]
        leaving = []
        for c in range(C):
            for w in arriving[c]:
                w.wait()
            ac = a[c]                                                      # (G, my planes of group c, g's y chunk, rest)
            if nd == 3:
                ac = ac.reshape((G, mc, n1 // G, self.rest[0]))
            axes[1].chunked_(ac, True)
            for d in range(2, nd - 1):
                axes[d].inverse_(ac, d + 1)
            rows(ac, out=ac)
            for d in range(nd - 2, 1, -1):
                axes[d].forward_(ac, d + 1)
            axes[1].chunked_(ac, False)
            leaving.append(self._exchange_lists([bv[g, c * mc:(c + 1) * mc] for g in range(G)], [a[c, g] for g in range(G)]))
        for ws in leaving:
            for w in ws:
                w.wait()
        return axes[0].forward_(b, 0, out=out)


def nls_slab_ops(k_axes: Sequence[torch.Tensor], gamma: float = 2.0, group=None) -> Tuple[torch.Tensor, Callable, SlabFFT]:
    """Slab-decomposed N-D cubic NLS (BASELINE cfg 5): returns this rank's ``lin_op`` block
    (spectral layout), the distributed ``nl_func`` and the transform object.

    u_t = i lap(u) + i gamma |u|^2 u;  L = -i |k|^2,  N(u^) = i gamma F{|f|^2 f},  f = F^-1{u^}.
    Pass ``group`` also to the solver so that the error norms are reduced globally:
        lin, nl, fft = nls_slab_ops([k, k, k], group=g);  ETD35(lin, nl, config, group=g)
    """
    shape = [int(k.shape[0]) for k in k_axes]
    fft = SlabFFT(shape, group)
    nd = len(shape)
    k2 = 0
    for d, k in enumerate(k_axes):
        view = [1] * nd
        view[d] = shape[d]
        k2 = k2 + (k.to(torch.float64) ** 2).reshape(view)
    lin_global = -1j * k2.to(torch.complex128)
    lin_op = fft.spec_slice(lin_global.expand(shape))

    own = None
    if lin_op.is_cuda:
        from . import _abi
        from .models import AxisFFT, RowNL, _pow2_in_range
        if _pow2_in_range(shape[-1]) and all(AxisFFT.supported(s) for s in shape[:-1]):
            cols = {s: AxisFFT(s, lin_op.device) for s in set(shape[:-1])}
            own = (RowNL(_abi.MODEL_NLS_FFT, shape[-1], None, gamma, lin_op.device), [cols[s] for s in shape[:-1]])

    def nl_func(uf: torch.Tensor, out=None) -> torch.Tensor:
        if own is not None and uf.is_contiguous():
            return fft.fused_nl(uf, own[0], own[1], out=out)
        f = fft.inverse(uf).contiguous()
        if f.is_cuda:
            from .models import pointwise_
            from . import _abi
            return fft.forward(pointwise_(_abi.MODEL_NLS_FFT, f, gamma))       # one pointwise kernel
        f2 = f.real ** 2 + f.imag ** 2                                         # gloo / CPU tests
        return 1j * gamma * fft.forward(f2 * f)

    nl_func.supports_out = True
    return lin_op, nl_func, fft
