"""GPU tests of RosyParameterization::integ() on the device (mk_igm_create_from_mesh / mk_igm_scatter, csrc/integ.cuh): the gradient stencil
G2 (integ.ipp:9-28), rawF2Vec (integ.ipp:30-43), X2F = (G2 * fullx2Nfunc * UFull).pruned() (integ.ipp:72, ir_init.h:173) and the result scatter
nfn / cfn (integ.ipp:88-98), against the host restatement of the same lines (tests/hostglue, which builds them with host SpGEMMs as the
reference does)."""
import numpy as np
import pytest

import oracle
from hostglue import Session
from metriko_b200 import api

pytestmark = pytest.mark.gpu


def _field_and_glue(ctx, s, seed):
    fi = s.field_inputs()
    F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
    z = F.solveSmallestEig(oracle.random_u0(s.dims["nF"], seed), 50)
    s.field_post(z, 4, "curl")
    s.igm_setup(0.02)
    return F, s.igm_inputs(), s.igm_mesh_inputs()


def _ints(o):
    return dict(zip([int(i) for i in o["fixedIdcs"]], [int(v) for v in o["fixedVals"]]))


@pytest.mark.parametrize("maker,seed,rs", [(lambda: Session().make_sphere(20, 3, bump=0.08), 0, False),
                                           (lambda: Session().make_torus(48, 32, bump=0.1), 0, True),
                                           (lambda: Session().make_sphere(8, 6, bump=0.1, phase=1.3), 4, False)])
def test_integ_on_device_matches_host(ctx, maker, seed, rs):
    s = maker()
    F, inp, m = _field_and_glue(ctx, s, seed)
    N = inp["N"]
    Xh = inp["X2F"].tocsr()
    # (1) host F2: X2F, A0 and the rounding result against the host-assembled operator
    prob = api.IgmProblem.from_mesh(ctx, N, m["cutF"], m["cutPos"], m["fullx2Nfunc"], inp["UFull"], inp["length"], F2=inp["F2"])
    Xd, F2d = prob.x2f()
    scale = abs(Xh).max()
    assert abs(Xd - Xh).max() <= 1e-13 * scale                      # entries: same sums in a different association order
    assert np.array_equal(F2d, np.asarray(inp["F2"], np.float64))
    Xl = (Xh * float(inp["length"])).tocsr()
    A = Xl.T @ Xl
    assert abs(prob.matrix() - A).max() <= 1e-12 * abs(A).max()
    out = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, rs)
    ref = api.iterative_rounding(ctx, inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], inp["length"], inp["UFull"], Xh,
                                 N, inp["n"], inp["nF"], True, rs, inp["F2"])
    assert out["success"] and ref["success"] and _ints(out) == _ints(ref) and out["n_solves"] == ref["n_solves"]
    assert np.abs(out["fullx"] - ref["fullx"]).max() < 1e-7
    # (2) result scatter on the device (fullx still resident) against the host restatement of integ.ipp:88-98
    nfn, cfn = prob.scatter()
    fin = s.igm_finish(out["fullx"])
    assert np.abs(nfn.reshape(-1) - fin["nfn"]).max() <= 1e-12 * max(1.0, np.abs(fin["nfn"]).max())
    assert np.abs(cfn.reshape(-1) - fin["cfn"]).max() <= 1e-12 * max(1.0, np.abs(fin["cfn"]).max())
    nfn2, _ = prob.scatter(out["fullx"], want_cfn=False)            # explicit host fullx gives the same
    assert np.array_equal(nfn2, nfn)
    prob.close()
    # (3) F2 formed on the device from the eigen-solve result still resident in the field handle + the combing turns
    prob2 = api.IgmProblem.from_mesh(ctx, N, m["cutF"], m["cutPos"], m["fullx2Nfunc"], inp["UFull"], inp["length"], field=F, comb_turns=m["comb_turns"])
    _, F2f = prob2.x2f()
    assert np.abs(F2f - inp["F2"]).max() <= 1e-13
    out2 = prob2.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, rs)
    assert out2["success"] and _ints(out2) == _ints(ref) and np.abs(out2["fullx"] - ref["fullx"]).max() < 1e-7
    prob2.close()


def test_from_mesh_argument_checks(ctx):
    s = Session().make_sphere(20, 2, bump=0.08)
    F, inp, m = _field_and_glue(ctx, s, 3)
    with pytest.raises(api.MkError):   # neither F2 nor (field, comb_turns)
        api.IgmProblem.from_mesh(ctx, inp["N"], m["cutF"], m["cutPos"], m["fullx2Nfunc"], inp["UFull"], inp["length"])
    prob = api.IgmProblem(ctx, inp["X2F"], inp["UFull"], inp["F2"], inp["length"])
    with pytest.raises(api.MkError):   # scatter needs the corner table: mk_igm_create_from_mesh handles only
        prob.N, prob.nF, prob.cut_nV = inp["N"], inp["nF"], len(m["cutPos"])
        prob.scatter(np.zeros(inp["UFull"].shape[0]))
