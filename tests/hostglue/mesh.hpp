// Host-side mesh stage: synthetic mesh generators and the flat half-edge arrays of the reference's
// Hmesh (reference include/metriko/core/hmesh/hmesh.h:313-569). This is NOT on the GPU hot path and
// NOT part of the drop-in: it restates the reference's unchanged host stage so that tests and the
// benchmark can produce hot-path inputs without Eigen/libigl (absent from this image).
// Index conventions follow SURVEY.md Appendix A.1/A.2 (igl::edge_topology + dcel numbering).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace hg {

struct V3 {
    double x = 0, y = 0, z = 0;
    V3() = default;
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    V3 operator+(const V3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    V3 operator-(const V3& o) const { return {x - o.x, y - o.y, z - o.z}; }
    V3 operator*(double s) const { return {x * s, y * s, z * s}; }
    V3 operator/(double s) const { return {x / s, y / s, z / s}; }
    double dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
    V3 cross(const V3& o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
    double norm() const { return std::sqrt(dot(*this)); }
    V3 normalized() const { double n = norm(); return {x / n, y / n, z / n}; }
};

struct RawMesh {
    std::vector<V3> V;
    std::vector<std::array<int, 3>> F;
};

// ----------------------------------------------------------------------------------------------
// generators (synthetic inputs of SURVEY.md §8d)
// ----------------------------------------------------------------------------------------------

// Torus with nu x nv quads, each split in two triangles; tube radius modulated to break symmetry.
inline RawMesh make_torus(int nu, int nv, double R, double r, double bump, double phase) {
    RawMesh m;
    m.V.resize((size_t)nu * nv);
    const double twoPi = 2 * M_PI;
    for (int i = 0; i < nu; i++) {
        for (int j = 0; j < nv; j++) {
            double th = twoPi * i / nu, ph = twoPi * j / nv;
            double rr = r * (1 + bump * std::sin(3 * ph + phase) * std::cos(2 * th - 0.5 * phase));
            m.V[(size_t)i * nv + j] = V3((R + rr * std::cos(ph)) * std::cos(th), (R + rr * std::cos(ph)) * std::sin(th),
                                         rr * std::sin(ph));
        }
    }
    m.F.reserve((size_t)2 * nu * nv);
    for (int i = 0; i < nu; i++) {
        for (int j = 0; j < nv; j++) {
            int i1 = (i + 1) % nu, j1 = (j + 1) % nv;
            int a = i * nv + j, b = i1 * nv + j, c = i1 * nv + j1, d = i * nv + j1;
            m.F.push_back({a, b, c});
            m.F.push_back({a, c, d});
        }
    }
    return m;
}

inline RawMesh base_polyhedron(int kind) {
    RawMesh m;
    if (kind == 4) {  // tetrahedron
        m.V = {V3(1, 1, 1), V3(1, -1, -1), V3(-1, 1, -1), V3(-1, -1, 1)};
        m.F = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
    } else if (kind == 8) {  // octahedron
        m.V = {V3(1, 0, 0), V3(-1, 0, 0), V3(0, 1, 0), V3(0, -1, 0), V3(0, 0, 1), V3(0, 0, -1)};
        m.F = {{0, 2, 4}, {2, 1, 4}, {1, 3, 4}, {3, 0, 4}, {2, 0, 5}, {1, 2, 5}, {3, 1, 5}, {0, 3, 5}};
    } else if (kind == 20) {  // icosahedron
        const double t = (1 + std::sqrt(5.0)) / 2;
        m.V = {V3(-1, t, 0), V3(1, t, 0), V3(-1, -t, 0), V3(1, -t, 0), V3(0, -1, t), V3(0, 1, t),
               V3(0, -1, -t), V3(0, 1, -t), V3(t, 0, -1), V3(t, 0, 1), V3(-t, 0, -1), V3(-t, 0, 1)};
        m.F = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
               {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
               {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    } else {
        throw std::runtime_error("base_polyhedron: kind must be 4, 8 or 20");
    }
    for (auto& p : m.V) p = p.normalized();
    // make sure faces are outward oriented
    for (auto& f : m.F) {
        V3 a = m.V[f[0]], b = m.V[f[1]], c = m.V[f[2]];
        if ((b - a).cross(c - a).dot(a + b + c) < 0) std::swap(f[1], f[2]);
    }
    return m;
}

// 1-to-4 midpoint subdivision, new vertices pushed to the unit sphere. Child faces of f are 4f..4f+3,
// so the face order stays hierarchical (good gather locality for the face-based field operator).
inline RawMesh subdivide_sphere(const RawMesh& in) {
    RawMesh out;
    const size_t nF = in.F.size();
    struct EK { int a, b; size_t slot; };
    std::vector<EK> keys(nF * 3);
    for (size_t f = 0; f < nF; f++)
        for (int j = 0; j < 3; j++) {
            int a = in.F[f][j], b = in.F[f][(j + 1) % 3];
            if (a > b) std::swap(a, b);
            keys[f * 3 + j] = {a, b, f * 3 + j};
        }
    std::sort(keys.begin(), keys.end(), [](const EK& p, const EK& q) { return p.a != q.a ? p.a < q.a : p.b < q.b; });
    out.V = in.V;
    std::vector<int> mid(nF * 3);
    for (size_t i = 0; i < keys.size();) {
        size_t j = i;
        int id = (int)out.V.size();
        out.V.push_back(((in.V[keys[i].a] + in.V[keys[i].b]) * 0.5).normalized());
        while (j < keys.size() && keys[j].a == keys[i].a && keys[j].b == keys[i].b) mid[keys[j++].slot] = id;
        i = j;
    }
    out.F.resize(nF * 4);
    for (size_t f = 0; f < nF; f++) {
        int v0 = in.F[f][0], v1 = in.F[f][1], v2 = in.F[f][2];
        int m01 = mid[f * 3 + 0], m12 = mid[f * 3 + 1], m20 = mid[f * 3 + 2];
        out.F[4 * f + 0] = {v0, m01, m20};
        out.F[4 * f + 1] = {m01, v1, m12};
        out.F[4 * f + 2] = {m20, m12, v2};
        out.F[4 * f + 3] = {m01, m12, m20};
    }
    return out;
}

// Sphere from a platonic solid subdivided `levels` times with a smooth low-frequency radial bump
// (a round sphere has a degenerate smallest eigenspace, SURVEY.md §7 "hard parts").
inline RawMesh make_sphere(int kind, int levels, double bump, double phase) {
    RawMesh m = base_polyhedron(kind);
    for (int l = 0; l < levels; l++) m = subdivide_sphere(m);
    for (auto& p : m.V) {
        double s = 1.0 + bump * (std::sin(2.0 * p.x + 0.3 + phase) * std::cos(1.5 * p.y - 0.2) +
                                 0.6 * std::sin(1.7 * p.z + 1.1 * p.x + 0.5 * phase) +
                                 0.4 * std::cos(2.3 * p.y * p.z - phase));
        p = p * s;
    }
    return m;
}

// ----------------------------------------------------------------------------------------------
// Hmesh flat arrays (reference hmesh.h:98-146)
// ----------------------------------------------------------------------------------------------
struct Hmesh {
    int nV = 0, nF = 0, nE = 0, nH = 0, nC = 0, euler = 0;
    std::vector<V3> pos;
    std::vector<std::array<int, 3>> idx;
    std::vector<std::array<int, 2>> edge2vert, edge2face;
    std::vector<std::array<int, 3>> face2edge;
    std::vector<int> vert2half, edge2half, face2half, crnr2half;
    std::vector<int> next, prev, twin, tail, head, edge, face, crnr;
    std::vector<V3> faceBasisX, faceBasisY, faceNormal;
    std::vector<double> faceArea, heArgOnFace, angleDefect;
    std::vector<double> dihedralArg, heLen;  // per half: dihedral angle (hmesh.h:508-519), |vec(h)| (hmesh.h:215)

    V3 vec(int h) const { return pos[head[h]] - pos[tail[h]]; }
    bool canonical(int h) const { return edge2half[edge[h]] == h; }
    // Vert::adjHalfs stepping (hmesh.h:158-161)
    int vnext_ccw(int h) const { return twin[prev[h]]; }
    int vnext_cw(int h) const { return next[twin[h]]; }
};

// Closed manifold meshes only (README.md:27 of the reference: boundaries unsupported).
inline Hmesh build_hmesh(const RawMesh& rm) {
    Hmesh m;
    m.pos = rm.V;
    m.idx = rm.F;
    m.nV = (int)rm.V.size();
    m.nF = (int)rm.F.size();
    const int nF = m.nF;
    // --- igl::edge_topology restated: edges = unique (vmin,vmax) pairs sorted
    //     lexicographically; edge2face(e,0) = the face in which the edge runs vmin -> vmax.
    struct ET { int a, b, f, j; };
    std::vector<ET> ett((size_t)nF * 3);
    for (int f = 0; f < nF; f++)
        for (int j = 0; j < 3; j++) {
            int a = rm.F[f][j], b = rm.F[f][(j + 1) % 3];
            if (a > b) std::swap(a, b);
            ett[(size_t)f * 3 + j] = {a, b, f, j};
        }
    std::sort(ett.begin(), ett.end(), [](const ET& p, const ET& q) {
        if (p.a != q.a) return p.a < q.a;
        if (p.b != q.b) return p.b < q.b;
        if (p.f != q.f) return p.f < q.f;
        return p.j < q.j;
    });
    m.face2edge.assign(nF, {-1, -1, -1});
    for (size_t i = 0; i < ett.size();) {
        if (i + 1 >= ett.size() || ett[i + 1].a != ett[i].a || ett[i + 1].b != ett[i].b)
            throw std::runtime_error("build_hmesh: mesh has a boundary edge (closed meshes only)");
        if (i + 2 < ett.size() && ett[i + 2].a == ett[i].a && ett[i + 2].b == ett[i].b)
            throw std::runtime_error("build_hmesh: non-manifold edge");
        int e = (int)m.edge2vert.size();
        m.edge2vert.push_back({ett[i].a, ett[i].b});
        int f0 = ett[i].f, f1 = ett[i + 1].f;
        m.face2edge[f0][ett[i].j] = e;
        m.face2edge[f1][ett[i + 1].j] = e;
        bool flip = true;
        for (int j = 0; j < 3; j++)
            if (rm.F[f0][j] == ett[i].a && rm.F[f0][(j + 1) % 3] == ett[i].b) flip = false;
        if (flip) std::swap(f0, f1);
        m.edge2face.push_back({f0, f1});
        i += 2;
    }
    m.nE = (int)m.edge2vert.size();
    m.nH = 2 * m.nE;
    m.nC = 3 * nF;
    m.euler = m.nV - m.nE + m.nF;
    const int nH = m.nH;
    // --- dcel (hmesh.h:240-311): half 2e in edge2face(e,0) with tail vmin, half 2e+1 its twin.
    m.vert2half.assign(m.nV, -1);
    m.edge2half.assign(m.nE, -1);
    m.face2half.assign(nF, -1);
    m.crnr2half.assign(m.nC, -1);
    m.next.assign(nH, -1); m.prev.assign(nH, -1); m.twin.assign(nH, -1);
    m.tail.assign(nH, -1); m.head.assign(nH, -1); m.edge.assign(nH, -1);
    m.face.assign(nH, -1); m.crnr.assign(nH, -1);
    std::vector<std::array<int, 3>> FH(nF, {-1, -1, -1});
    for (int e = 0; e < m.nE; e++) {
        for (int k = 0; k < 2; k++) {
            int h = 2 * e + k, f = m.edge2face[e][k];
            m.edge[h] = e;
            m.tail[h] = m.edge2vert[e][k];
            m.vert2half[m.edge2vert[e][k]] = h;  // last write wins (hmesh.h:272-281)
            m.face[h] = f;
            m.twin[h] = 2 * e + (1 - k);
            for (int j = 0; j < 3; j++)
                if (m.face2edge[f][j] == e) FH[f][j] = h;
        }
        m.edge2half[e] = 2 * e;
    }
    for (int f = 0; f < nF; f++) {
        m.face2half[f] = FH[f][0];
        for (int j = 0; j < 3; j++) {
            m.next[FH[f][j]] = FH[f][(j + 1) % 3];
            m.prev[FH[f][(j + 1) % 3]] = FH[f][j];
        }
    }
    for (int h = 0; h < nH; h++) m.head[h] = m.tail[m.next[h]];
    // corners: crnr[h] = the corner opposite to h (hmesh.h:430-441)
    for (int f = 0; f < nF; f++)
        for (int it = 0; it < 3; it++) {
            int vid = rm.F[f][it], cid = 3 * f + it;
            int hc = m.face2half[f];
            for (int h : {hc, m.next[hc], m.prev[hc]})
                if (m.tail[h] != vid && m.head[h] != vid) { m.crnr2half[cid] = h; m.crnr[h] = cid; }
        }
    // --- geometry caches (hmesh.h:461-489, 552-567)
    m.faceBasisX.resize(nF); m.faceBasisY.resize(nF); m.faceNormal.resize(nF);
    m.faceArea.resize(nF); m.heArgOnFace.assign(nH, 0.0); m.angleDefect.resize(m.nV);
    for (int f = 0; f < nF; f++) {
        int h0 = m.face2half[f];
        V3 x = m.vec(h0);
        V3 t = m.vec(m.prev[h0]) * -1.0;
        V3 n = x.cross(t);
        m.faceBasisX[f] = x.normalized();
        m.faceBasisY[f] = (x.cross(n) * -1.0).normalized();
        m.faceNormal[f] = n.normalized();
        m.faceArea[f] = n.norm() * 0.5;
        double sum = 0;
        int h = h0;
        for (int k = 0; k < 3; k++) {
            V3 v1 = m.vec(h).normalized();
            V3 v2 = m.vec(m.prev[h]).normalized();
            if ((v1 - m.faceBasisX[f]).norm() > 1e-10) sum += std::acos(v1.dot(v2));
            m.heArgOnFace[h] = sum;
            h = m.next[h];
        }
    }
    for (int v = 0; v < m.nV; v++) {
        double sum = 0;
        int h0 = m.vert2half[v], h = h0;
        do {
            V3 v1 = m.vec(h).normalized();
            V3 v2 = m.vec(m.twin[m.prev[h]]).normalized();
            sum += std::acos(std::min(1.0, std::max(-1.0, v1.dot(v2))));
            h = m.vnext_ccw(h);
        } while (h != h0);
        m.angleDefect[v] = 2 * M_PI - sum;
    }
    // dihedral angle per half (hmesh.h:508-519; closed meshes: no boundary edges) and half length (hmesh.h:215)
    m.dihedralArg.resize(nH);
    m.heLen.resize(nH);
    for (int h = 0; h < nH; h++) {
        const V3 n1 = m.faceNormal[m.face[h]], n2 = m.faceNormal[m.face[m.twin[h]]];
        const V3 e = m.vec(h);
        const double len = e.norm();
        const V3 v = e * (1.0 / len);
        m.heLen[h] = len;
        m.dihedralArg[h] = std::atan2(v.dot(n1.cross(n2)), n1.dot(n2));
    }
    return m;
}

}  // namespace hg
