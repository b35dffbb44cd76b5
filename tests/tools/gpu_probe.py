"""Exploratory GPU run: field + IGM parity against the oracle on small meshes, SpMV bandwidth on a large one."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
from hostglue import Session
from metriko_b200 import api, capi
import oracle
from oracle.field import field_angle_error

def log(*a):
    print(*a, flush=True)

ctx = capi.Context(0)
levels = int(os.environ.get("PROBE_LEVELS", "4"))
s = Session().make_sphere(20, levels, bump=0.08)
fi = s.field_inputs(); nF = s.dims['nF']
log("mesh", s.dims)
u0 = oracle.random_u0(nF, 0)
L = oracle.connection_laplacian(fi['twin'], fi['face'], fi['heArgOnFace'], nF, 4)
M = oracle.galerkin_mass_matrix(fi['faceArea'])
t = time.time(); z_ref = oracle.solve_smallest_eig(L, M, u0, 50); log("oracle eig s", time.time() - t)
F = api.FaceRosyField(ctx, fi['face2half'], fi['next'], fi['twin'], fi['face'], fi['heArgOnFace'], fi['faceArea'], 4)
Ld, md = F.operator()
log("assembly diff", abs(Ld - L.tocsr()).max(), abs(md - fi['faceArea']).max())
x = u0.copy(); log("spmv diff", np.abs(F.spmv(x) - L @ x).max())
for pre in (0, 1):
    t = time.time()
    try:
        z = F.solveSmallestEig(u0, 50, precond=pre)
        err = field_angle_error(z, z_ref, 4)
        log(f"precond={pre} wall {time.time()-t:.3f}s stats {F.stats} max angle err {err.max():.3e}")
    except Exception as e:
        log("precond", pre, "FAILED", repr(e), F.stats)
# csc path
try:
    t = time.time(); z2, st = api.solveSmallestEig(ctx, L, M, 50, u0, return_stats=True)
    log("csc path", time.time() - t, st, "err", field_angle_error(z2, z_ref, 4).max())
except Exception as e:
    log("csc path FAILED", repr(e))
# IGM
s.field_post(z_ref, 4, 'curl'); d = s.igm_setup(0.02); inp = s.igm_inputs(); log("igm dims", d)
t = time.time()
ref = oracle.iterative_rounding(inp['fixedIdcs'], inp['fixedVals'], inp['singularIdcs'], inp['integerIdcs'], inp['length'], inp['UFull'], inp['G2p'], inp['N'], inp['n'], inp['nF'], True, False, inp['F2'], ni_mode='closed')
log("oracle igm s", time.time() - t, "solves", ref['n_solves'])
for pre in (0, 1):
    try:
        t = time.time()
        prob = api.IgmProblem(ctx, inp['X2F'], inp['UFull'], inp['F2'], inp['length'], precond=pre)
        A0 = prob.matrix(); X = (inp['X2F'] * inp['length']).tocsr(); log("A0 diff", abs(A0 - (X.T @ X)).max())
        xs, xf = prob.poisson(inp['fixedIdcs'], inp['fixedVals'])
        log(f"precond={pre} poisson", prob.stats, "dx", np.abs(xf - ref['x0_full']).max())
        out = prob.round(inp['fixedIdcs'], inp['fixedVals'], inp['singularIdcs'], inp['integerIdcs'], True, False)
        same = dict(zip(out['fixedIdcs'], out['fixedVals'])) == dict(zip(ref['fixedIdcs'], ref['fixedVals']))
        log(f"precond={pre} round wall {time.time()-t:.3f}s success {out['success']} same ints {same} solves {out['n_solves']} dx {np.abs(out['fullx']-ref['fullx']).max():.3e} stats {out['stats']}")
        fin = s.igm_finish(out['fullx']); log({k: v for k, v in fin.items() if k not in ('cfn', 'nfn')})
    except Exception as e:
        import traceback; traceback.print_exc(); log("igm precond", pre, "FAILED", repr(e))
# bandwidth on a large mesh
big = int(os.environ.get("PROBE_BIG", "9"))
s2 = Session().make_sphere(8, big, bump=0.08); fi2 = s2.field_inputs(); nF2 = s2.dims['nF']
log("big mesh", s2.dims)
F2 = api.FaceRosyField(ctx, fi2['face2half'], fi2['next'], fi2['twin'], fi2['face'], fi2['heArgOnFace'], fi2['faceArea'], 4)
for fused in (0, 1):
    ms = F2.spmv_bench(50, bool(fused))
    log(f"L spmv fused={fused}: {ms*1e3:.1f} us, algorithmic 116 B/row -> {116*nF2/ms/1e6:.1f} GB/s, moved 92 B/row -> {92*nF2/ms/1e6:.1f} GB/s")
u02 = oracle.random_u0(nF2, 0)
for pre in (1,):
    t = time.time()
    try:
        z = F2.solveSmallestEig(u02, int(os.environ.get("PROBE_ITERS", "50")), precond=pre)
        log(f"big precond={pre} wall {time.time()-t:.3f}s stats {F2.stats}")
    except Exception as e:
        log("big precond", pre, "FAILED", repr(e), F2.stats)
