// C-ABI of the host-glue library (libmetriko_hostglue.so). One "session" owns a mesh, the stages
// derived from it and a bag of named flat arrays that Python reads/writes through ctypes.
// Nothing here touches the GPU; see mesh.hpp / field_glue.hpp / igm_glue.hpp for what each stage restates.
#include <array>
#include <cstdio>
#include <memory>
#include <string>

#include "bag.hpp"
#include "mesh.hpp"
#include "field_glue.hpp"
#include "igm_glue.hpp"

namespace hg {
struct Session {
    Bag bag;
    Hmesh mesh;
    bool has_mesh = false;
    FieldStage field;
    IgmStage igm;
    std::string err;
};

static void export_mesh(Session& s) {
    const Hmesh& m = s.mesh;
    Bag& b = s.bag;
    std::vector<double> pos((size_t)m.nV * 3);
    for (int i = 0; i < m.nV; i++) { pos[3 * i] = m.pos[i].x; pos[3 * i + 1] = m.pos[i].y; pos[3 * i + 2] = m.pos[i].z; }
    b.put("pos", pos);
    b.put("idx", m.idx);
    b.put("edge2vert", m.edge2vert);
    b.put("edge2face", m.edge2face);
    b.put("face2edge", m.face2edge);
    b.put("vert2half", m.vert2half);
    b.put("face2half", m.face2half);
    b.put("next", m.next); b.put("prev", m.prev); b.put("twin", m.twin);
    b.put("tail", m.tail); b.put("head", m.head); b.put("edge", m.edge);
    b.put("face", m.face); b.put("crnr", m.crnr);
    b.put("faceArea", m.faceArea);
    b.put("heArgOnFace", m.heArgOnFace);
    b.put("angleDefect", m.angleDefect);
    b.put("dihedralArg", m.dihedralArg);
    b.put("heLen", m.heLen);
    auto put3 = [&](const char* name, const std::vector<V3>& v) {
        std::vector<double> o(v.size() * 3);
        for (size_t i = 0; i < v.size(); i++) { o[3 * i] = v[i].x; o[3 * i + 1] = v[i].y; o[3 * i + 2] = v[i].z; }
        b.put(name, o);
    };
    put3("faceBasisX", m.faceBasisX); put3("faceBasisY", m.faceBasisY); put3("faceNormal", m.faceNormal);
    std::vector<int> dims = {m.nV, m.nF, m.nE, m.nH, m.euler};
    b.put("dims", dims);
}
}  // namespace hg

using hg::Session;

#define HG_TRY(s, ...)                                     \
    try { __VA_ARGS__; return 0; }                         \
    catch (const std::exception& e) { (s)->err = e.what(); return -1; } \
    catch (...) { (s)->err = "unknown error"; return -1; }

extern "C" {

void* hg_session_new() { return new Session(); }
void hg_session_free(void* p) { delete (Session*)p; }
const char* hg_last_error(void* p) { return ((Session*)p)->err.c_str(); }

int hg_bag_put(void* p, const char* name, int dtype, const void* data, int64_t n) {
    Session* s = (Session*)p;
    HG_TRY(s, s->bag.put_raw(name, dtype, data, n));
}
int hg_bag_get(void* p, const char* name, const void** data, int64_t* n, int* dtype) {
    Session* s = (Session*)p;
    HG_TRY(s, { const hg::Arr& a = s->bag.get(name); *data = a.data.data(); *n = a.n; *dtype = a.dtype; });
}
int hg_bag_has(void* p, const char* name) { return ((Session*)p)->bag.arrs.count(name) ? 1 : 0; }

// kind: 0 = torus (a = nu, b = nv), 4/8/20 = subdivided tetra/octa/icosahedron (a = levels)
int hg_make_mesh(void* p, int kind, int a, int b, double bump, double phase) {
    Session* s = (Session*)p;
    HG_TRY(s, {
        hg::RawMesh rm = kind == 0 ? hg::make_torus(a, b, 1.0, 0.4, bump, phase) : hg::make_sphere(kind, a, bump, phase);
        s->mesh = hg::build_hmesh(rm);
        s->has_mesh = true;
        hg::export_mesh(*s);
    });
}

// mesh from caller-provided V (nV x 3 doubles) and F (nF x 3 ints)
int hg_set_mesh(void* p, const double* V, int nV, const int* F, int nF) {
    Session* s = (Session*)p;
    HG_TRY(s, {
        hg::RawMesh rm;
        rm.V.resize(nV);
        rm.F.resize(nF);
        for (int i = 0; i < nV; i++) rm.V[i] = hg::V3(V[3 * i], V[3 * i + 1], V[3 * i + 2]);
        for (int i = 0; i < nF; i++) rm.F[i] = {F[3 * i], F[3 * i + 1], F[3 * i + 2]};
        s->mesh = hg::build_hmesh(rm);
        s->has_mesh = true;
        hg::export_mesh(*s);
    });
}

// Field post-processing (frf.h:37-43, 100-202): needs bag["compressed"] (complex nF). Writes
// field, matching, effort, singular, seam, combed field/matching, extrinsic field.
int hg_field_post(void* p, int N, int matching_type) {
    Session* s = (Session*)p;
    HG_TRY(s, {
        if (!s->has_mesh) throw std::runtime_error("no mesh");
        hg::run_field_stage(s->mesh, s->bag, N, matching_type, s->field);
    });
}

// IGM host stage (param.h:126-152, setup.ipp, integ.ipp:4-43, ir_init.h:98-154): writes the
// iterative_rounding inputs (G2' = G2*fullx2Nfunc, Cfull, UFull, X2F, F2, idcs, length, ...).
int hg_igm_setup(void* p, double gridscale) {
    Session* s = (Session*)p;
    HG_TRY(s, {
        if (!s->has_mesh) throw std::runtime_error("no mesh");
        hg::run_igm_stage(s->mesh, s->field, s->bag, gridscale, s->igm);
    });
}

// integ.ipp:88-98 result scatter + invariants (orientation count, deviation): needs bag["fullx"].
int hg_igm_finish(void* p) {
    Session* s = (Session*)p;
    HG_TRY(s, hg::run_igm_finish(s->mesh, s->field, s->bag, s->igm));
}

}  // extern "C"
