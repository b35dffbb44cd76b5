"""GPU twin of tests/test_oracle.py::test_barrier_rounding_against_golden: iterative_rounding with localInjectivity = true (ir.h:57-112;
injective_barrier.h:19-109) through mk_igm_round_injective, against the oracle's golden vectors (tests/golden/*_barrier.npz) -- identical
integers, fullx within 1e-6 (post_checking's own tolerance), the NI iterate [x; f] of the barrier LM, and the example_qgp flags
(seamless = false, localInjectivity = true; example_qgp.cpp:26-29 -- BASELINE configs[0]) on the reference's icosphere.obj."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, load_golden
from golden.make_golden_large import RECIPES, make_session
from metriko_b200 import api

pytestmark = pytest.mark.gpu


def _ints(o):
    return dict(zip([int(i) for i in o["fixedIdcs"]], [int(v) for v in o["fixedVals"]]))


@pytest.mark.parametrize("name", ["sphere_ico3", "torus_24x16"])
def test_barrier_rounding_against_golden_gpu(ctx, name):
    g = load_golden(name)
    gb = dict(np.load(os.path.join(GOLDEN, name + "_barrier.npz")))
    prob = api.IgmProblem(ctx, g["X2F"], g["UFull"], g["F2"], float(g["length"]))
    # NI stage alone (seamless = false returns the NI iterate, ir.h:86-88)
    ni = prob.round(g["fixedIdcs"], g["fixedVals"], g["singularIdcs"], g["integerIdcs"], False, bool(g["round_seams"]), localInjectivity=True)
    assert ni["success"] and ni["fixedIdcs"] == []
    assert np.abs(ni["fullx"] - g["UFull"] @ gb["x_ni"]).max() < 1e-6
    # NI + rounding loop
    out = prob.round(g["fixedIdcs"], g["fixedVals"], g["singularIdcs"], g["integerIdcs"], True, bool(g["round_seams"]), localInjectivity=True)
    assert out["success"]
    assert _ints(out) == dict(zip(gb["out_fixedIdcs"].tolist(), gb["out_fixedVals"].tolist()))       # bit-exact integers
    assert out["n_solves"] == int(gb["n_solves"])
    assert np.abs(out["fullx"] - gb["fullx"]).max() < 1e-6
    # the barrier did its job: no field pair is flipped and the smallest volume ratio is the oracle's
    from oracle.barrier import InjectiveBarrier
    B = InjectiveBarrier(4, len(g["F2"]) // 8, g["F2"])
    X = g["X2F"] * float(g["length"])
    x_small = np.linalg.lstsq(g["UFull"].toarray(), out["fullx"], rcond=None)[0] if g["UFull"].shape[0] < 4000 else None
    if x_small is not None:
        r = B.ratio(X @ x_small)[0]
        assert r.min() > 0 and abs(r.min() - float(gb["min_ratio"])) < 1e-5
    prob.close()


def test_example_qgp_flags_on_the_reference_icosphere(ctx):
    """BASELINE configs[0]: example/models/icosphere.obj at scale 0.03 with example_qgp's flags seamless = false, localInjectivity = true
    (example_qgp.cpp:26-29): field on the device, integ() on the device from the cut-mesh arrays, the barrier NI LM with the reference's
    iter_cb (stop as soon as the map is locally injective, integ.ipp:46-63) -- against the oracle on the same inputs."""
    recipe, seed, rs, scale, _ = RECIPES["icosphere"]
    s = make_session(recipe)
    fi = s.field_inputs()
    F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
    z = F.solveSmallestEig(oracle.random_u0(s.dims["nF"], seed), 50)
    s.field_post(z, 4, "curl")
    s.igm_setup(scale)
    inp, m = s.igm_inputs(), s.igm_mesh_inputs()
    F2N, cutF = m["fullx2Nfunc"].tocsr(), m["cutF"]

    def injective(fullx):   # locally_injectivity_check (integ.ipp:46-63)
        nfn = (F2N @ fullx).reshape(-1, 4)
        a, b, c = nfn[cutF[:, 0], :2], nfn[cutF[:, 1], :2], nfn[cutF[:, 2], :2]
        d1, d2 = b - a, c - a
        return not np.any(d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0] < 0)

    ref = oracle.iterative_rounding_x2f(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], inp["length"], inp["UFull"],
                                        inp["X2F"], 4, 2, False, rs, inp["F2"], localInjectivity=True, iter_cb=injective)
    prob = api.IgmProblem.from_mesh(ctx, 4, m["cutF"], m["cutPos"], m["fullx2Nfunc"], inp["UFull"], inp["length"], field=F, comb_turns=m["comb_turns"])
    out = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], False, rs, localInjectivity=True, use_iter_cb=True)
    assert out["success"] and np.abs(out["fullx"] - ref["fullx"]).max() < 1e-6
    assert injective(out["fullx"]) == injective(ref["fullx"])
    nfn, cfn = prob.scatter()
    fin = s.igm_finish(out["fullx"])
    assert np.abs(nfn.reshape(-1) - fin["nfn"]).max() < 1e-10
    prob.close()
