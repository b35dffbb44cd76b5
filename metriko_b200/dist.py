"""Partitioned multi-GPU face operator (one process per GPU): Python driver over the `mk_dist_*` / `mk_dfield_*` C-ABI.

torch.distributed is only the plumbing that hands rank 0's NCCL id to the other ranks; the halo exchange and the
all-reduces of the distributed PCG run inside libmetriko_b200.so on its own NCCL communicator (csrc/dist.cuh).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import Context, Stats, as_c128, as_f64, as_i32, default_opts, ptr
from .partition import HaloPlan, make_halo_plan

ID_BYTES = 128


def face_neighbours(face2half, next_, twin, face) -> np.ndarray:
    """[3, nF] global neighbour face of every half of every face, in the half order face2half, next, next(next)
    (the column pattern of connectionLaplacian, frf.h:62-71). Pure index bookkeeping (host)."""
    h = np.asarray(face2half, np.int64).copy()
    nxt, tw, fc = np.asarray(next_), np.asarray(twin), np.asarray(face)
    out = np.empty((3, len(h)), np.int64)
    for j in range(3):
        out[j] = fc[tw[h]]
        h = nxt[h]
    return out


class DistContext:
    """mk_dist: NCCL communicator of the partitioned mode on top of a capi.Context."""

    def __init__(self, ctx: Context, rank: int, world: int, bcast=None):
        """bcast(bytes_or_None) -> bytes: broadcasts rank 0's id to every rank (default: torch.distributed)."""
        self.ctx, self.rank, self.world = ctx, rank, world
        L = ctx.L
        buf = C.create_string_buffer(ID_BYTES)
        if world > 1:
            if rank == 0:
                ctx.check(L.mk_dist_unique_id(buf, ID_BYTES))
            if bcast is None:
                import torch
                import torch.distributed as dist
                t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
                if dist.get_backend() == "nccl":
                    t = t.cuda(ctx.device)
                dist.broadcast(t, 0)
                raw = bytes(t.cpu().numpy().tobytes())
            else:
                raw = bcast(buf.raw if rank == 0 else None)
            buf = C.create_string_buffer(raw, ID_BYTES)
        self.h = C.c_void_p()
        ctx.check(L.mk_dist_create(ctx.h, rank, world, buf, ID_BYTES, C.byref(self.h)))

    def close(self):
        if self.h:
            self.ctx.L.mk_dist_destroy(self.h)
            self.h = C.c_void_p()


class DistFaceField:
    """This rank's rows of the connection Laplacian + mass (frf.h:46-84) with the halo plan of metriko_b200.partition."""

    def __init__(self, dctx: DistContext, plan: HaloPlan, val_local, mass_local):
        self.d, self.plan = dctx, plan
        ctx = dctx.ctx
        self.ctx = ctx
        self.h = C.c_void_p()
        a = [as_i32(plan.nbr_local), as_c128(val_local), as_f64(mass_local), as_i32(plan.send_counts), as_i32(plan.recv_counts),
             as_i32(plan.send_idx)]
        assert a[1].shape == (3, plan.n_local) and a[2].shape == (plan.n_local,)
        ctx.check(ctx.L.mk_dfield_create(dctx.h, plan.n_local, plan.n_interior, plan.n_halo, *[ptr(v) for v in a], C.byref(self.h)))
        self.stats = None

    def set_global(self, nbr_global, val_global):
        """Hands every rank the global stencil (nbr / val [3, nF]) so that the solves use the GLOBAL multigrid hierarchy (level 0
        distributed, restricted residual all-reduced, coarse levels replicated) instead of block-Jacobi. Collective."""
        nbr, val = as_i32(nbr_global), as_c128(val_global)
        assert nbr.shape == val.shape and nbr.shape[0] == 3
        l2g = as_i32(self.plan.owned_global())
        self.ctx.check(self.ctx.L.mk_dfield_set_global(self.h, nbr.shape[1], ptr(nbr), ptr(val), ptr(l2g)))

    @classmethod
    def from_mesh(cls, dctx: DistContext, face2half, next_, twin, face, heArgOnFace, faceArea, N: int = 4, ranges=None,
                  global_hierarchy: bool = True):
        """Partition + local assembly. The neighbour pattern is index bookkeeping on the host; the values
        -transport^N = -(-1)^N e^{iN(farg(h) - farg(twin h))} (frf.h:52-71) of the owned rows are evaluated with numpy here
        (set-up, O(rows/world)); the single-GPU path assembles them on the device (k_field_assemble)."""
        nbr = face_neighbours(face2half, next_, twin, face)
        plan = make_halo_plan(nbr, dctx.rank, dctx.world, ranges)
        rows = plan.owned_global()
        farg = np.asarray(heArgOnFace, np.float64)
        tw, nxt = np.asarray(twin), np.asarray(next_)
        h = np.asarray(face2half, np.int64)
        if not global_hierarchy:
            h = h[rows]
        val = np.empty((3, len(h)), np.complex128)
        sign = -1.0 if (N & 1) else 1.0
        for j in range(3):
            val[j] = -sign * np.exp(1j * N * (farg[h] - farg[tw[h]]))
            h = nxt[h]
        self = cls(dctx, plan, val[:, rows] if global_hierarchy else val, np.asarray(faceArea, np.float64)[rows])
        if global_hierarchy:
            self.set_global(nbr, val)
        return self

    def spmv(self, x_local):
        x = as_c128(x_local)
        y = np.empty_like(x)
        self.ctx.check(self.ctx.L.mk_dfield_spmv(self.h, ptr(x), ptr(y)))
        return y

    def spmv_bench(self, reps: int = 50, overlap: bool = True) -> float:
        ms = C.c_double()
        self.ctx.check(self.ctx.L.mk_dfield_spmv_bench(self.h, reps, int(overlap), C.byref(ms)))
        return ms.value

    def pcg(self, b_local, x0_local=None, **opts):
        b = as_c128(b_local)
        x0 = as_c128(x0_local) if x0_local is not None else None
        x = np.empty_like(b)
        st = Stats()
        o = default_opts(**opts)
        rc = self.ctx.L.mk_dfield_pcg(self.h, ptr(b), ptr(x0), C.byref(o), ptr(x), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc, allow=(-3,))
        return x

    def pcg_dev(self, b_dev_ptr: int, x_dev_ptr: int, **opts):
        """The same solve with the owned rows of b and x (start vector in, solution out) resident in device memory (raw pointers)."""
        st = Stats()
        o = default_opts(**opts)
        rc = self.ctx.L.mk_dfield_pcg_dev(self.h, C.c_void_p(b_dev_ptr), C.c_void_p(x_dev_ptr), C.byref(o), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc, allow=(-3,))

    def solveSmallestEig(self, u0_local, nIter: int = 50, **opts):
        u0 = as_c128(u0_local)
        x = np.empty_like(u0)
        st = Stats()
        o = default_opts(**opts)
        rc = self.ctx.L.mk_dfield_solve(self.h, ptr(u0), nIter, C.byref(o), ptr(x), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc, allow=(-3,))
        return x

    def close(self):
        if self.h:
            self.ctx.L.mk_dfield_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
