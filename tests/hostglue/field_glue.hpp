// Host stage between the field eigen-solve and the IGM solve: restates the reference's serial glue in
// include/metriko/core/vectorfield/face_rosy_field.h (decode :37-43, connection :109-117, curl/principal
// matching :138-202, singularities :100-107, seam :268-360, combing :207-249, extrinsic field :251-264).
// OUT OF SCOPE of the GPU path -- it exists so hot-path inputs can be generated here.
// libigl tie-breaks (cut graph, dijkstra) are restated from their published algorithms (parity unpinned).
#pragma once
#include <complex>
#include <deque>
#include <queue>
#include <set>

#include "bag.hpp"
#include "mesh.hpp"

namespace hg {

using cplx = std::complex<double>;

struct FieldStage {
    int N = 4;
    std::vector<cplx> field;        // nF x N (row major), raw field
    std::vector<cplx> connection;   // nE
    std::vector<int> matching;      // nE raw matching
    std::vector<double> effort;     // nE
    std::vector<int> singular;      // nV
    std::vector<char> seam;         // nE
    std::vector<cplx> cfield;       // combed field nF x N
    std::vector<int> cmatching;     // combed matching nE
    std::vector<double> ext;        // nF x 3N extrinsic combed field
};

// frf.h:37-43
inline void decode_field(const std::vector<cplx>& z, int N, std::vector<cplx>& field) {
    const size_t nF = z.size();
    field.resize(nF * N);
    const cplx r = std::polar(1.0, 2 * M_PI / N);
    for (size_t f = 0; f < nF; f++) {
        cplx base = std::pow(std::polar(1.0, std::arg(z[f])), 1.0 / N);
        for (int i = 0; i < N; i++) field[f * N + i] = base * std::pow(r, (double)i);
    }
}

// igl::triangle_triangle_adjacency equivalents from the half-edge arrays
inline void tt_adjacency(const Hmesh& m, std::vector<std::array<int, 3>>& TT, std::vector<std::array<int, 3>>& TTi) {
    TT.assign(m.nF, {-1, -1, -1});
    TTi.assign(m.nF, {-1, -1, -1});
    for (int f = 0; f < m.nF; f++)
        for (int j = 0; j < 3; j++) {
            int e = m.face2edge[f][j];
            int g = m.edge2face[e][0] == f ? m.edge2face[e][1] : m.edge2face[e][0];
            TT[f][j] = g;
            for (int k = 0; k < 3; k++)
                if (m.face2edge[g][k] == e) TTi[f][j] = k;
        }
}

// libigl cut_mesh_from_singularities with an all-zero mismatch (frf.h:279) -> spanning-tree
// complement retracted to the 2g generator loops; then each singular vertex is joined by a
// unit-weight shortest path (frf.h:268-327).
inline void compute_seam(const Hmesh& m, const std::vector<int>& singular, std::vector<char>& seam) {
    const int nF = m.nF, nV = m.nV;
    std::vector<std::array<int, 3>> TT, TTi;
    tt_adjacency(m, TT, TTi);
    std::vector<std::array<char, 3>> cuts(nF, {1, 1, 1});
    {
        std::vector<char> visited(nF, 0);
        std::deque<int> d;
        for (int s = 0; s < nF; s++) {
            if (visited[s]) continue;
            visited[s] = 1;
            d.push_back(s);
            while (!d.empty()) {
                int f = d.front();
                d.pop_front();
                for (int k = 0; k < 3; k++) {
                    int g = TT[f][k], j = TTi[f][k];
                    if (!visited[g]) {
                        cuts[f][k] = 0;
                        cuts[g][j] = 0;
                        visited[g] = 1;
                        d.push_back(g);
                    }
                }
            }
        }
        // retract valence-1 ends
        std::vector<int> val(nV, 0);
        for (int f = 0; f < nF; f++)
            for (int k = 0; k < 3; k++)
                if (f < TT[f][k] && cuts[f][k]) { val[m.idx[f][k]]++; val[m.idx[f][(k + 1) % 3]]++; }
        bool guard = true;
        while (guard) {
            guard = false;
            for (int f = 0; f < nF; f++)
                for (int k = 0; k < 3; k++)
                    if (f < TT[f][k] && cuts[f][k]) {
                        int x0 = m.idx[f][k], x1 = m.idx[f][(k + 1) % 3];
                        if (val[x0] == 1 || val[x1] == 1) {
                            cuts[f][k] = 0;
                            cuts[TT[f][k]][TTi[f][k]] = 0;
                            val[x0]--;
                            val[x1]--;
                            guard = true;
                        }
                    }
        }
    }
    std::vector<char> in_cut(nV, 0);
    bool any_cut = false;
    for (int f = 0; f < nF; f++)
        for (int k = 0; k < 3; k++)
            if (cuts[f][k]) { in_cut[m.idx[f][k]] = 1; any_cut = true; }  // frf.h:281-285: F(i,j) only
    (void)any_cut;
    // vertex adjacency (igl::adjacency_list) via outgoing halfs
    std::vector<int> svids;
    for (int v = 0; v < nV; v++)
        if (singular[v] != 0) svids.push_back(v);
    std::vector<double> dist(nV);
    std::vector<int> prevv(nV);
    for (int s : svids) {
        // igl::dijkstra(source, targets, VV, min_distance, previous): set-ordered by (dist, id)
        std::fill(dist.begin(), dist.end(), 1e300);
        std::fill(prevv.begin(), prevv.end(), -1);
        std::set<std::pair<double, int>> q;
        dist[s] = 0;
        q.insert({0.0, s});
        int found = -1;
        while (!q.empty()) {
            auto [du, u] = *q.begin();
            q.erase(q.begin());
            if (in_cut[u]) { found = u; break; }
            int h0 = m.vert2half[u], h = h0;
            do {
                int v = m.head[h];
                double dn = du + 1.0;
                if (dn < dist[v]) {
                    q.erase({dist[v], v});
                    dist[v] = dn;
                    prevv[v] = u;
                    q.insert({dn, v});
                }
                h = m.vnext_ccw(h);
            } while (h != h0);
        }
        std::vector<int> path;
        if (found == -1) path.push_back(s);
        else
            for (int v = found; v != -1; v = prevv[v]) path.push_back(v);
        for (int v : path) in_cut[v] = 1;
        for (size_t i = 0; i + 1 < path.size(); i++) {
            int v0 = path[i], v1 = path[i + 1];
            // find the half v0->v1, mark both sides
            int h0 = m.vert2half[v0], h = h0;
            do {
                if (m.head[h] == v1) break;
                h = m.vnext_ccw(h);
            } while (h != h0);
            if (m.head[h] != v1) throw std::runtime_error("compute_seam: path edge not found");
            for (int hh : {h, m.twin[h]}) {
                int f = m.face[hh];
                for (int k = 0; k < 3; k++)
                    if (m.face2edge[f][k] == m.edge[hh]) cuts[f][k] = 1;
            }
        }
    }
    seam.assign(m.nE, 0);
    for (int f = 0; f < nF; f++)
        for (int k = 0; k < 3; k++)
            if (cuts[f][k]) seam[m.face2edge[f][k]] = 1;
}

inline void run_field_stage(const Hmesh& m, Bag& bag, int N, int matching_type, FieldStage& fs) {
    fs.N = N;
    std::vector<cplx> z = bag.get_c128("compressed");
    if ((int)z.size() != m.nF) throw std::runtime_error("field_post: compressed has wrong size");
    decode_field(z, N, fs.field);
    const auto& fld = fs.field;
    auto F = [&](int f, int i) { return fld[(size_t)f * N + i]; };
    // connection (frf.h:109-117)
    fs.connection.resize(m.nE);
    for (int e = 0; e < m.nE; e++) {
        V3 v = m.vec(2 * e).normalized();
        int f0 = m.edge2face[e][0], f1 = m.edge2face[e][1];
        cplx ef(v.dot(m.faceBasisX[f0]), v.dot(m.faceBasisY[f0]));
        cplx eg(v.dot(m.faceBasisX[f1]), v.dot(m.faceBasisY[f1]));
        fs.connection[e] = eg / ef;
    }
    fs.matching.assign(m.nE, -1);
    fs.effort.assign(m.nE, 0.0);
    for (int e = 0; e < m.nE; e++) {
        int f0 = m.edge2face[e][0], f1 = m.edge2face[e][1];
        cplx conn = fs.connection[e];
        if (matching_type == 1) {  // Curl (frf.h:169-202)
            int iM = 0;
            double mn = 32767000.;
            V3 v = m.vec(2 * e).normalized();
            for (int i = 0; i < N; i++) {
                double curl = 0;
                for (int j = 0; j < N; j++) {
                    cplx c0 = F(f0, j), c1 = F(f1, (i + j) % N);
                    V3 v0 = m.faceBasisX[f0] * c0.real() + m.faceBasisY[f0] * c0.imag();
                    V3 v1 = m.faceBasisX[f1] * c1.real() + m.faceBasisY[f1] * c1.imag();
                    double d = v.dot(v1 - v0);
                    curl += d * d;
                }
                if (curl < mn) { iM = i; mn = curl; }
            }
            fs.matching[e] = iM;
            cplx coef(1, 0);
            for (int i = 0; i < N; i++) coef *= F(f1, (i + iM) % N) / (F(f0, i) * conn);
            fs.effort[e] = std::arg(coef);
        } else {  // Principal (frf.h:138-167)
            double minRot = 1000.;
            int offset = 0;
            cplx coef(1, 0);
            cplx transport0 = conn * F(f0, 0);
            for (int i = 0; i < N; i++) {
                cplx jf = F(f0, i), jg = F(f1, i);
                coef *= jg / (conn * jf);
                double r = std::arg(jg / transport0);
                if (std::fabs(r) < std::fabs(minRot)) { offset = i; minRot = r; }
            }
            fs.effort[e] = std::arg(coef);
            double residuals = 0;
            for (int i = 0; i < N; i++) residuals += std::arg(F(f1, (i + offset) % N) / (conn * F(f0, i)));
            fs.matching[e] = offset - (int)std::round((residuals - fs.effort[e]) / (2 * M_PI));
        }
    }
    // singularities (frf.h:100-107)
    fs.singular.assign(m.nV, 1);
    for (int v = 0; v < m.nV; v++) {
        double sum = m.angleDefect[v] * N;
        int h0 = m.vert2half[v], h = h0;
        do {
            sum += (m.canonical(h) ? -1 : 1) * fs.effort[m.edge[h]];
            h = m.vnext_ccw(h);
        } while (h != h0);
        fs.singular[v] = (int)std::round(sum / (2 * M_PI));
    }
    compute_seam(m, fs.singular, fs.seam);
    // combing (frf.h:207-249)
    fs.cfield.assign(fld.size(), cplx(0, 0));
    std::vector<int> turns(m.nF, 0);
    std::vector<char> visit(m.nF, 0);
    std::queue<std::pair<int, int>> Q;
    Q.emplace(0, 0);
    do {
        auto [i, mm] = Q.front();
        Q.pop();
        if (visit[i]) continue;
        visit[i] = 1;
        turns[i] = mm;
        int r = N - mm;
        for (int k = 0; k < r; k++) fs.cfield[(size_t)i * N + k] = F(i, mm + k);
        for (int k = 0; k < mm; k++) fs.cfield[(size_t)i * N + r + k] = F(i, k);
        int h0 = m.face2half[i], h = h0;
        for (int k = 0; k < 3; k++, h = m.next[h]) {
            int e = m.edge[h];
            bool b = m.edge2face[e][0] == i;
            int fn = b ? m.edge2face[e][1] : m.edge2face[e][0];
            int n = (b ? 1 : -1) * fs.matching[e];
            n = (n + mm + 1000 * N) % N;
            if (!visit[fn] && !fs.seam[e]) Q.emplace(fn, n);
        }
    } while (!Q.empty());
    fs.cmatching.assign(m.nE, -1);
    for (int e = 0; e < m.nE; e++)
        fs.cmatching[e] = (turns[m.edge2face[e][0]] - turns[m.edge2face[e][1]] + fs.matching[e] + 1000 * N) % N;
    // extrinsic field of the combed field (frf.h:251-264)
    fs.ext.assign((size_t)m.nF * 3 * N, 0.0);
    for (int f = 0; f < m.nF; f++)
        for (int i = 0; i < N; i++) {
            cplx c = fs.cfield[(size_t)f * N + i];
            V3 d = (m.faceBasisX[f] * c.real() + m.faceBasisY[f] * c.imag()).normalized();
            fs.ext[(size_t)f * 3 * N + 3 * i + 0] = d.x;
            fs.ext[(size_t)f * 3 * N + 3 * i + 1] = d.y;
            fs.ext[(size_t)f * 3 * N + 3 * i + 2] = d.z;
        }
    bag.put("field", fs.field);
    bag.put("matching", fs.matching);
    bag.put("effort", fs.effort);
    bag.put("singular", fs.singular);
    std::vector<int> seam_i(fs.seam.begin(), fs.seam.end());
    bag.put("seam", seam_i);
    bag.put("cfield", fs.cfield);
    bag.put("comb_turns", turns);  // per-face turn of the combing BFS (frf.h:207-249): cfield(f, k) = field(f, (k + turn) % N)
    bag.put("cmatching", fs.cmatching);
    bag.put("ext", fs.ext);
    int nvisited = 0;
    for (char c : visit) nvisited += c;
    bag.put_scalar_i("comb_visited", nvisited);
}

}  // namespace hg
