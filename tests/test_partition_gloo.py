"""world_size 2 / 3 `gloo` tests (CPU) of the host logic behind the partitioned multi-GPU path: the halo plan
(metriko_b200/partition.py) and the sequence of steps of the distributed SpMV / CG (pack -> exchange -> interior rows ->
boundary rows; one all-reduce per reduction) on a numpy emulation of the device kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from metriko_b200.partition import block_ranges, make_halo_plan, morton_order


def ell3(g, N=4):
    """numpy restatement of k_field_assemble (frf.h:46-76): nbr[3, nF], val[3, nF], diagonal 3 + 1e-9."""
    f2h, nxt, twin, face, farg = g["face2half"], g["next"], g["twin"], g["face"], g["heArgOnFace"]
    nF = len(f2h)
    nbr = np.empty((3, nF), np.int64)
    val = np.empty((3, nF), np.complex128)
    h = f2h.copy()
    for j in range(3):
        t = twin[h]
        nbr[j] = face[t]
        val[j] = -((-1.0) ** N) * np.exp(1j * N * (farg[h] - farg[t]))
        h = nxt[h]
    return nbr, val


def global_spmv(nbr, val, x):
    return (3.0 + 1e-9) * x + (val * x[nbr]).sum(0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def exchange(plan, x_ext):
    """halo exchange of the emulation: isend / irecv per peer (what ncclSend / ncclRecv do in csrc/dist.cuh)"""
    send = x_ext[plan.send_idx]                      # pack
    reqs, soff, roff, bufs = [], 0, plan.n_local, []
    for p in range(plan.world):
        ns, nr = int(plan.send_counts[p]), int(plan.recv_counts[p])
        if ns:
            t = torch.from_numpy(np.ascontiguousarray(send[soff:soff + ns]).view(np.float64).copy())
            reqs.append(dist.isend(t, p))
            bufs.append(t)
        if nr:
            t = torch.empty(2 * nr, dtype=torch.float64)
            reqs.append(dist.irecv(t, p))
            bufs.append((t, roff, nr))
        soff += ns
        roff += nr
    for r in reqs:
        r.wait()
    for b in bufs:
        if isinstance(b, tuple):
            t, o, n = b
            x_ext[o:o + n] = t.numpy().view(np.complex128)


def dist_spmv(plan, val_loc, x_ext):
    """y = L x on the owned rows: interior rows need no halo (they overlap the exchange on the GPU), boundary rows after it"""
    y = np.empty(plan.n_local, np.complex128)
    ni = plan.n_interior
    y[:ni] = (3.0 + 1e-9) * x_ext[:ni] + (val_loc[:, :ni] * x_ext[plan.nbr_local[:, :ni]]).sum(0)
    exchange(plan, x_ext)
    y[ni:] = (3.0 + 1e-9) * x_ext[ni:plan.n_local] + (val_loc[:, ni:] * x_ext[plan.nbr_local[:, ni:]]).sum(0)
    return y


def allreduce(*vals):
    t = torch.tensor(vals, dtype=torch.float64)
    dist.all_reduce(t)
    return t.tolist()


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(name)
        nbr, val = ell3(g)
        nF = nbr.shape[1]
        plan = make_halo_plan(nbr, rank, world)
        rows = plan.owned_global()
        val_loc = val[:, rows]
        # --- plan invariants
        assert plan.n_local == block_ranges(nF, world)[rank + 1] - block_ranges(nF, world)[rank]
        assert (plan.nbr_local[:, :plan.n_interior] < plan.n_local).all()
        assert (plan.nbr_local[:, plan.n_interior:] >= plan.n_local).any(axis=0).all()
        counts = torch.from_numpy(plan.send_counts.copy())
        allc = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        for p in range(world):       # what p sends to me is what I expect from p
            assert int(allc[p][rank]) == int(plan.recv_counts[p])
        # --- distributed SpMV equals the global one
        rng = np.random.default_rng(7)
        x = rng.standard_normal(nF) + 1j * rng.standard_normal(nF)
        x_ext = np.zeros(plan.n_local + plan.n_halo, np.complex128)
        x_ext[:plan.n_local] = plan.scatter_from_global(x)
        y = dist_spmv(plan, val_loc, x_ext)
        y_ref = global_spmv(nbr, val, x)[rows]
        assert np.abs(y - y_ref).max() < 1e-13
        assert np.array_equal(x_ext[plan.n_local:], x[plan.halo_global])
        # --- distributed CG (two all-reduces per iteration, as the CUDA driver) equals a direct solve
        b = (g["faceArea"] * g["u0"])[rows]
        xs = np.zeros(plan.n_local + plan.n_halo, np.complex128)
        r = b.copy()
        p_ext = np.zeros_like(xs)
        p_ext[:plan.n_local] = r
        (rho, bb) = allreduce(np.vdot(r, r).real, np.vdot(b, b).real)
        it = 0
        while rho > 1e-24 * bb and it < 5000:
            qv = dist_spmv(plan, val_loc, p_ext)
            (pq,) = allreduce(np.vdot(p_ext[:plan.n_local], qv).real)
            a = rho / pq
            xs[:plan.n_local] += a * p_ext[:plan.n_local]
            r -= a * qv
            (rho1,) = allreduce(np.vdot(r, r).real)
            p_ext[:plan.n_local] = r + (rho1 / rho) * p_ext[:plan.n_local]
            rho = rho1
            it += 1
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        nFr = np.arange(nF)
        L = sp.csr_matrix((np.concatenate([np.full(nF, 3.0 + 1e-9, np.complex128), val.ravel()]),
                           (np.concatenate([nFr, nFr, nFr, nFr]), np.concatenate([nFr, nbr.ravel()]))), shape=(nF, nF))
        x_ref = spla.spsolve(L.tocsc(), g["faceArea"] * g["u0"])
        assert np.abs(xs[:plan.n_local] - x_ref[rows]).max() < 1e-9 * np.abs(x_ref).max()
        q.put((rank, "ok", it))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "sphere_ico3"), (3, "torus_24x16")])
def test_partitioned_spmv_and_cg_gloo(world, name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"


def test_halo_plan_single_rank_and_morton():
    g = load_golden("sphere_ico3")
    nbr, _ = ell3(g)
    plan = make_halo_plan(nbr, 0, 1)
    assert plan.n_halo == 0 and plan.n_interior == plan.n_local == nbr.shape[1] and len(plan.send_idx) == 0
    # a Z-order renumbering keeps most neighbours inside a contiguous block: the cut of a 4-way block partition shrinks
    rng = np.random.default_rng(0)
    shuffle = rng.permutation(nbr.shape[1])
    inv = np.empty_like(shuffle)
    inv[shuffle] = np.arange(len(shuffle))
    nbr_shuf = inv[nbr[:, shuffle]]                     # same graph, random row order
    cut_shuf = sum(make_halo_plan(nbr_shuf, r, 4).n_halo for r in range(4))
    # centroids are not in the fixture: use a spectral-free stand-in, the original (locality-preserving) order
    cut_orig = sum(make_halo_plan(nbr, r, 4).n_halo for r in range(4))
    assert cut_orig < 0.5 * cut_shuf
    pts = rng.standard_normal((1000, 3))
    order = morton_order(pts)
    assert sorted(order.tolist()) == list(range(1000))
    d_sorted = np.linalg.norm(np.diff(pts[order], axis=0), axis=1).mean()
    d_rand = np.linalg.norm(np.diff(pts, axis=0), axis=1).mean()
    assert d_sorted < 0.6 * d_rand


# ---------------------------------------------------------------------------------------------------------------------
# The global-hierarchy preconditioner of the partitioned solves (csrc/dist.cuh DistAmgPrecond): level 0 distributed, restricted
# residual all-reduced, coarse levels replicated. Emulated here with numpy + gloo on the hierarchy of the host mirror; it must
# reproduce the single-process V-cycle exactly (same PCG iteration count), which is what makes the multi-GPU iteration counts
# those of one GPU rather than of block-Jacobi.
# ---------------------------------------------------------------------------------------------------------------------
def _cheb(apply_A, dinv, lmax, b, x, zero, deg=2, ratio=8.0):
    lmin = lmax / ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    r = b.copy() if zero else b - apply_A(x)
    d = dinv * r / theta
    if zero:
        x = np.zeros_like(b)
    for _ in range(deg):
        x = x + d
        r = r - apply_A(d)
        rho1 = 1.0 / (2 * sigma - rho)
        d = rho1 * rho * d + 2 * rho1 / delta * dinv * r
        rho = rho1
    return x, r


def _worker_global_amg(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        from amg_mirror import VCycle, pcg
        from metriko_b200 import capi
        g = load_golden(name)
        nbr, val = ell3(g)
        nF = nbr.shape[1]
        plan = make_halo_plan(nbr, rank, world)
        rows = plan.owned_global()
        val_loc = val[:, rows]
        nFr = np.arange(nF)
        L = sp.csr_matrix((np.concatenate([np.full(nF, 3.0 + 1e-9, np.complex128), val.ravel()]),
                           (np.concatenate([nFr, nFr, nFr, nFr]), np.concatenate([nFr, nbr.ravel()]))), shape=(nF, nF))
        H = capi.amg_host_hierarchy(L)                 # every rank builds the same (deterministic) global hierarchy
        assert len(H) >= 2
        P0 = H[0][1].tocsr()[rows]                     # my rows of P_0, local order
        R0 = P0.conj().T.tocsr()
        dinv0, lmax0 = H[0][3][rows], H[0][2]
        coarse = VCycle(H[1:])                         # replicated levels >= 1

        def apply_A(v_loc):
            x_ext = np.zeros(plan.n_local + plan.n_halo, np.complex128)
            x_ext[:plan.n_local] = v_loc
            return dist_spmv(plan, val_loc, x_ext)

        def precond(r_loc):
            x, r = _cheb(apply_A, dinv0, lmax0, r_loc, None, True)
            part = R0 @ r                              # partial restriction over my fine rows
            t = torch.from_numpy(np.ascontiguousarray(part).view(np.float64).copy())
            dist.all_reduce(t)                         # the one data all-reduce of the V-cycle
            x = x + P0 @ coarse(t.numpy().view(np.complex128))
            x, _ = _cheb(apply_A, dinv0, lmax0, r_loc, x, False)
            return x

        b = (g["faceArea"] * g["u0"])[rows]
        x = np.zeros(plan.n_local, np.complex128)
        r = b.copy()
        z = precond(r)
        p = z.copy()
        (rho, bb) = allreduce(np.vdot(r, z).real, np.vdot(b, b).real)
        it = 0
        while it < 200:
            (rr,) = allreduce(np.vdot(r, r).real)
            if rr <= 1e-20 * bb:
                break
            qv = apply_A(p)
            (pq,) = allreduce(np.vdot(p, qv).real)
            a = rho / pq
            x += a * p
            r -= a * qv
            z = precond(r)
            (rho1,) = allreduce(np.vdot(r, z).real)
            p = z + (rho1 / rho) * p
            rho = rho1
            it += 1
        # single-process reference with the same hierarchy
        bg = g["faceArea"] * g["u0"]
        x_ref, it_ref = pcg(L, bg, np.zeros_like(bg), VCycle(H))
        assert abs(it - it_ref) <= 1, (it, it_ref)
        assert np.abs(x - x_ref[rows]).max() < 1e-8 * np.abs(x_ref).max()
        q.put((rank, "ok", it))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_global_hierarchy_preconditioner_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker_global_amg, args=(r, world, port, "sphere_ico3", q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"
    assert res[0][2] == res[1][2] and res[0][2] < 40
