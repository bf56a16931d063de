// Host-side construction of the k-independent tables (index tables, quadrature, node-function
// recurrences, coupling coefficients).  Everything numeric is done in long double and rounded once.
//
// Replaces (for chain coordinate types a / ba / bba / b..ba):
//   ush.index_array_harmonics, ush.flatten_harmonics           (_biem.py:651,686,720,743,889,917)
//   the Gauss product rule inside ush.expand(..., n=n_end)      (_biem.py:627)
//   the harmonic triple integrals inside ush.harmonics_translation_coef (_biem.py:697)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <thread>

#include "plan.h"

typedef long double ld;
static const ld PI_L = 3.14159265358979323846264338327950288L;

// ---- index tables ---------------------------------------------------------------------------------
static void build_index(int s_ndim, int n_end, std::vector<int32_t>& out) {
    out.clear();
    std::vector<int> ms;
    for (int m = 0; m < n_end; ++m) ms.push_back(m);
    for (int m = -(n_end - 1); m < 0; ++m) ms.push_back(m);
    std::vector<int> cur(s_ndim, 0);
    // iterative DFS over n_0 >= n_1 >= ... >= |m|
    struct Rec {
        static void go(int depth, int upper, int s_ndim, const std::vector<int>& ms, std::vector<int>& cur,
                       std::vector<int32_t>& out) {
            if (depth == s_ndim - 1) {
                for (int m : ms)
                    if (std::abs(m) <= upper) {
                        for (int i = 0; i < s_ndim - 1; ++i) out.push_back(cur[i]);
                        out.push_back(m);
                    }
                return;
            }
            for (int v = 0; v <= upper; ++v) {
                cur[depth] = v;
                go(depth + 1, v, s_ndim, ms, cur, out);
            }
        }
    };
    Rec::go(0, n_end - 1, s_ndim, ms, cur, out);
}

// ---- ultraspherical node functions ----------------------------------------------------------------
// W_p = int_0^pi sin^p
static ld wallis(int p) {
    ld w = (p & 1) ? 2.0L : PI_L;
    for (int q = (p & 1) ? 3 : 2; q <= p; q += 2) w *= (ld)(q - 1) / (ld)q;
    return w;
}
// a_{n,l}^2 for a node with `desc` descendants
static ld a2coef(int n, int l, int desc) {
    ld hn = (ld)n + 0.5L * desc;
    return (ld)(n - l) * (ld)(n + l + desc - 1) / (4.0L * hn * (hn - 1.0L));
}
// f_{n,l}(theta) for all l <= n < Lb at x = cos, st = sin:  out[n*Lb + l]
static void node_functions(int desc, int Lb, ld x, ld st, std::vector<ld>& out) {
    out.assign((size_t)Lb * Lb, 0.0L);
    ld sl = 1.0L;
    for (int l = 0; l < Lb; ++l) {
        ld f0 = sl / sqrtl(wallis(2 * l + desc));
        out[(size_t)l * Lb + l] = f0;
        ld fm2 = 0.0L, fm1 = f0;
        for (int n = l + 1; n < Lb; ++n) {
            ld an = sqrtl(a2coef(n, l, desc));
            ld anm1 = (n - 1 > l) ? sqrtl(a2coef(n - 1, l, desc)) : 0.0L;
            ld f = (x * fm1 - anm1 * fm2) / an;
            out[(size_t)n * Lb + l] = f;
            fm2 = fm1;
            fm1 = f;
        }
        sl *= st;
    }
}

// ---- Gauss-Gegenbauer quadrature: weight (1-x^2)^a, a = (desc-1)/2 ------------------------------------
// orthonormal polynomial p_n and derivative via the recurrence x p_k = a_{k+1} p_{k+1} + a_k p_{k-1}
static void gauss_gegenbauer(int n, int desc, std::vector<ld>& xs, std::vector<ld>& ws) {
    xs.resize(n);
    ws.resize(n);
    std::vector<ld> a(n + 2, 0.0L);
    for (int k = 1; k <= n + 1; ++k) a[k] = sqrtl(a2coef(k, 0, desc));
    ld mu0 = wallis(desc);  // int (1-x^2)^{(desc-1)/2} dx = int sin^desc
    ld p0 = 1.0L / sqrtl(mu0);
    // symmetric tridiagonal eigenvalues by bisection (Sturm count) -- robust, n is small
    auto sturm = [&](ld lam) {
        int cnt = 0;
        ld q = -lam;
        if (q < 0) ++cnt;
        for (int k = 1; k < n; ++k) {
            ld qq = (q == 0.0L) ? 1e-4000L : q;
            q = -lam - a[k] * a[k] / qq;
            if (q < 0) ++cnt;
        }
        return cnt;  // number of eigenvalues < lam
    };
    for (int i = 0; i < n; ++i) {
        ld lo = -1.0L, hi = 1.0L;
        for (int it = 0; it < 80; ++it) {
            ld mid = 0.5L * (lo + hi);
            if (sturm(mid) > i) hi = mid; else lo = mid;
        }
        ld x = 0.5L * (lo + hi);
        // Newton polish on p_n
        for (int it = 0; it < 4; ++it) {
            ld pm1 = 0.0L, p = p0, dm1 = 0.0L, dp = 0.0L;
            for (int k = 0; k < n; ++k) {
                ld pn = (x * p - a[k] * pm1) / a[k + 1];
                ld dn = (p + x * dp - a[k] * dm1) / a[k + 1];
                pm1 = p; p = pn; dm1 = dp; dp = dn;
            }
            x -= p / dp;
        }
        // Christoffel weight 1 / sum_{k<n} p_k(x)^2
        ld pm1 = 0.0L, p = p0, sum = 0.0L;
        for (int k = 0; k < n; ++k) {
            sum += p * p;
            ld pn = (x * p - a[k] * pm1) / a[k + 1];
            pm1 = p; p = pn;
        }
        xs[i] = x;
        ws[i] = 1.0L / sum;
    }
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& h) {
    *dptr = nullptr;
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    if (cudaMalloc((void**)dptr, bytes) != cudaSuccess) return BHS_ERR_ALLOC;
    if (!h.empty() && cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
        return BHS_ERR_ALLOC;
    return BHS_OK;
}

int bhs_fill_WY(bhs_plan* p);             // harmonics.cu
int bhs_fill_planar_tables(bhs_plan* p);  // harmonics.cu

// ---- coupling table -------------------------------------------------------------------------------
struct Term {
    int idx;
    double coef;
};

static int build_coupling(bhs_plan* p) {
    const int d = p->d, s = p->s_ndim, L = p->n_end, L2 = p->L2, H = p->H;
    const ld cd = powl(2.0L * PI_L, 0.5L * d) * sqrtl(2.0L / PI_L);
    const ld inv_s2pi = 1.0L / sqrtl(2.0L * PI_L);
    const int32_t* tab = p->h_idx.data();
    // lookup for band-2 indices
    std::map<std::vector<int>, int> lookup2;
    for (int h = 0; h < p->H2; ++h) {
        std::vector<int> key(p->h_idx2.begin() + (size_t)h * s, p->h_idx2.begin() + (size_t)(h + 1) * s);
        lookup2[key] = h;
    }
    // node-function values at the triple-integral quadrature nodes
    const int nq = 2 * L + 1;
    std::vector<std::vector<ld>> F(p->n_bnodes);  // [node][(n*L2+l)*nq + q]
    std::vector<std::vector<ld>> W(p->n_bnodes);
    for (int i = 0; i < p->n_bnodes; ++i) {
        int desc = d - 2 - i;
        std::vector<ld> xs, ws, tmp;
        gauss_gegenbauer(nq, desc, xs, ws);
        W[i] = ws;
        F[i].resize((size_t)nq * L2 * L2);
        for (int q = 0; q < nq; ++q) {
            node_functions(desc, L2, xs[q], sqrtl(std::max((ld)0.0L, 1.0L - xs[q] * xs[q])), tmp);
            for (size_t e = 0; e < tmp.size(); ++e) F[i][e * nq + q] = tmp[e];
        }
    }
    auto triple = [&](int node, int np, int lp, int n, int l, int n2, int l2) -> ld {
        const ld* f = F[node].data();
        const ld* fa = f + ((size_t)np * L2 + lp) * nq;
        const ld* fb = f + ((size_t)n * L2 + l) * nq;
        const ld* fc = f + ((size_t)n2 * L2 + l2) * nq;
        const ld* w = W[node].data();
        ld acc = 0.0L;
        for (int q = 0; q < nq; ++q) acc += w[q] * fa[q] * fb[q] * fc[q];
        return acc;
    };

    // 2-D (Graf): exactly one term per entry with a constant coefficient -- the tiles are filled directly (the
    // generic path below would need H^2 small vectors, 47 M of them at the reference's largest run n_end = 3444).
    // The i^{|m|+|m2|-|mp|} factor is split: i^{|m2|} into SY, i^{|m|} into the row and (-i)^{|mp|} into the column factor.
    if (d == 2) {
        const int Lb2 = p->L2;  // band-2 ordering: m2 = 0..Lb2-1, then -(Lb2-1)..-1
        auto idx2 = [&](int m2) { return m2 >= 0 ? m2 : 2 * Lb2 - 1 + m2; };
        p->coupling_terms = (int64_t)H * H;
        p->tiles_r = (H + BHS_TILE_R - 1) / BHS_TILE_R;
        p->tiles_c = (H + BHS_TILE_C - 1) / BHS_TILE_C;
        const size_t ntile = (size_t)p->tiles_r * p->tiles_c;
        p->h_tiles.assign(ntile, bhs_tile_hdr());
        std::vector<double> coef(ntile * BHS_TILE_E, 0.0);
        std::vector<uint16_t> cidx(ntile * BHS_TILE_E, 0);
        p->max_nt = 1;
        p->max_sy_cnt = 1;
        for (int tr = 0; tr < p->tiles_r; ++tr)
            for (int tc = 0; tc < p->tiles_c; ++tc) {
                const size_t ti = (size_t)tr * p->tiles_c + tc;
                bhs_tile_hdr& hd = p->h_tiles[ti];
                int lo = 1 << 30, hi = -1;
                for (int r = 0; r < BHS_TILE_R; ++r)
                    for (int c = 0; c < BHS_TILE_C; ++c) {
                        int h = tr * BHS_TILE_R + r, hp = tc * BHS_TILE_C + c;
                        if (h >= H || hp >= H) continue;
                        int id = idx2(tab[hp] - tab[h]);
                        lo = std::min(lo, id);
                        hi = std::max(hi, id);
                    }
                if (hi < 0) { lo = 0; hi = 0; }
                if (hi - lo > 65535) return BHS_ERR_UNSUPPORTED;
                hd.nt = 1; hd.sy_lo = lo; hd.sy_cnt = hi - lo + 1; hd.pad = 0;
                p->max_sy_cnt = std::max(p->max_sy_cnt, hd.sy_cnt);
                hd.coef_off = (int64_t)(ti * BHS_TILE_E);
                hd.idx_off = (int64_t)(ti * BHS_TILE_E);
                for (int r = 0; r < BHS_TILE_R; ++r)
                    for (int c = 0; c < BHS_TILE_C; ++c) {
                        int h = tr * BHS_TILE_R + r, hp = tc * BHS_TILE_C + c;
                        if (h >= H || hp >= H) continue;
                        const size_t o = ti * BHS_TILE_E + (size_t)r * BHS_TILE_C + c;
                        coef[o] = (double)(cd * inv_s2pi);
                        cidx[o] = (uint16_t)(idx2(tab[hp] - tab[h]) - lo);
                    }
            }
        p->coupling_bytes = (int64_t)(coef.size() * sizeof(double) + cidx.size() * sizeof(uint16_t) +
                                      p->h_tiles.size() * sizeof(bhs_tile_hdr));
        int rc;
        if ((rc = upload(&p->d_coef, coef))) return rc;
        if ((rc = upload(&p->d_cidx, cidx))) return rc;
        if ((rc = upload(&p->d_tiles, p->h_tiles))) return rc;
        return BHS_OK;
    }
    // terms[(h_row, h'_col)] : row = h (harmonic of ball b), col = h' (harmonic of ball b')
    std::vector<std::vector<Term>> terms((size_t)H * H);
    // rows h are independent: split them over host threads (the triple integrals are long-double quadratures,
    // 2.5e9 multiply-adds at n_end = 39)
    if (d < 3) return BHS_ERR_UNSUPPORTED;
    const std::map<std::vector<int>, int>& lk2 = lookup2;  // read-only from here on
    auto row_work = [&](int h, int64_t& cnt) {
        if (d == 3) {
            int n = tab[h * 2], m = tab[h * 2 + 1];
            for (int hp = 0; hp < H; ++hp) {
                int np = tab[hp * 2], mp = tab[hp * 2 + 1];
                int m2 = mp - m;
                std::vector<Term>& tv = terms[(size_t)h * H + hp];
                for (int n2 = std::abs(n - np); n2 <= n + np; n2 += 2) {
                    if (n2 < std::abs(m2)) continue;
                    ld g = triple(0, np, std::abs(mp), n, std::abs(m), n2, std::abs(m2)) * inv_s2pi;
                    tv.push_back({lk2.at({n2, m2}), (double)(cd * g)});
                    ++cnt;
                }
            }
        } else if (d >= 5) {
            // chains of any depth: one Gegenbauer triple integral per b-node (node i has d-2-i descendants); the lower
            // indices of node i are the degrees chosen at node i+1 (|m| at the innermost b-node).  Enumerated from the
            // azimuth outwards with an explicit stack of (n''_i) choices.
            const int nb = s - 1;  // number of b-nodes
            const int32_t* ih = tab + (size_t)h * s;
            for (int hp = 0; hp < H; ++hp) {
                const int32_t* ip = tab + (size_t)hp * s;
                const int m2 = ip[s - 1] - ih[s - 1];
                std::vector<Term>& tv = terms[(size_t)h * H + hp];
                int n2[BHS_MAX_NODES + 1];
                ld gacc[BHS_MAX_NODES + 2];
                // depth-first over nodes i = nb-1 .. 0
                int i = nb - 1;
                gacc[nb] = 1.0L;
                n2[nb] = std::abs(m2);  // lower index of the innermost node
                int cur[BHS_MAX_NODES + 1];
                auto first = [&](int node) {
                    int lo = std::abs(ih[node] - ip[node]);
                    int low2 = n2[node + 1];
                    while (lo < low2) lo += 2;
                    return lo;
                };
                cur[i] = first(i);
                while (i < nb) {
                    if (cur[i] > ih[i] + ip[i]) {  // exhausted: back up
                        ++i;
                        if (i < nb) cur[i] += 2;
                        continue;
                    }
                    const int lowp = (i + 1 < nb) ? ip[i + 1] : std::abs(ip[s - 1]);
                    const int lowh = (i + 1 < nb) ? ih[i + 1] : std::abs(ih[s - 1]);
                    n2[i] = cur[i];
                    gacc[i] = gacc[i + 1] * triple(i, ip[i], lowp, ih[i], lowh, n2[i], n2[i + 1]);
                    if (i == 0) {
                        std::vector<int> key(s);
                        for (int q = 0; q < nb; ++q) key[q] = n2[q];
                        key[s - 1] = m2;
                        tv.push_back({lk2.at(key), (double)(cd * gacc[0] * inv_s2pi)});
                        ++cnt;
                        cur[0] += 2;
                    } else {
                        --i;
                        cur[i] = first(i);
                    }
                }
            }
        } else {
            int n = tab[h * 3], l = tab[h * 3 + 1], m = tab[h * 3 + 2];
            for (int hp = 0; hp < H; ++hp) {
                int np = tab[hp * 3], lp = tab[hp * 3 + 1], mp = tab[hp * 3 + 2];
                int m2 = mp - m;
                std::vector<Term>& tv = terms[(size_t)h * H + hp];
                for (int l2 = std::abs(l - lp); l2 <= l + lp; l2 += 2) {
                    if (l2 < std::abs(m2)) continue;
                    ld g1 = triple(1, lp, std::abs(mp), l, std::abs(m), l2, std::abs(m2));
                    for (int n2 = std::abs(n - np); n2 <= n + np; n2 += 2) {
                        if (n2 < l2) continue;
                        ld g = triple(0, np, lp, n, l, n2, l2) * g1 * inv_s2pi;
                        tv.push_back({lk2.at({n2, l2, m2}), (double)(cd * g)});
                        ++cnt;
                    }
                }
            }
        }
    };
    int64_t nterms = 0;
    {
        unsigned hw = std::thread::hardware_concurrency();
        int T = (int)std::min<unsigned>(hw ? hw : 1, 16);
        if ((int64_t)H * H < 4096) T = 1;
        std::vector<int64_t> cnts(T, 0);
        std::vector<std::thread> pool;
        for (int t = 1; t < T; ++t)
            pool.emplace_back([&, t]() { for (int h = t; h < H; h += T) row_work(h, cnts[t]); });
        for (int h = 0; h < H; h += T) row_work(h, cnts[0]);
        for (auto& th : pool) th.join();
        for (int64_t c : cnts) nterms += c;
    }
    p->coupling_terms = nterms;

    // tile-wise ELL layout
    p->tiles_r = (H + BHS_TILE_R - 1) / BHS_TILE_R;
    p->tiles_c = (H + BHS_TILE_C - 1) / BHS_TILE_C;
    p->h_tiles.assign((size_t)p->tiles_r * p->tiles_c, bhs_tile_hdr());
    std::vector<double> coef;
    std::vector<uint16_t> cidx;
    p->max_nt = 0;
    p->max_sy_cnt = 1;
    for (int tr = 0; tr < p->tiles_r; ++tr)
        for (int tc = 0; tc < p->tiles_c; ++tc) {
            bhs_tile_hdr& hd = p->h_tiles[(size_t)tr * p->tiles_c + tc];
            int nt = 0, lo = 1 << 30, hi = -1;
            for (int r = 0; r < BHS_TILE_R; ++r)
                for (int c = 0; c < BHS_TILE_C; ++c) {
                    int h = tr * BHS_TILE_R + r, hp = tc * BHS_TILE_C + c;
                    if (h >= H || hp >= H) continue;
                    const std::vector<Term>& tv = terms[(size_t)h * H + hp];
                    nt = std::max(nt, (int)tv.size());
                    for (const Term& t : tv) { lo = std::min(lo, t.idx); hi = std::max(hi, t.idx); }
                }
            if (hi < 0) { lo = 0; hi = 0; }
            hd.nt = nt;
            hd.sy_lo = lo;
            hd.sy_cnt = hi - lo + 1;
            p->max_sy_cnt = std::max(p->max_sy_cnt, hd.sy_cnt);
            hd.pad = 0;
            hd.coef_off = (int64_t)coef.size();
            hd.idx_off = (int64_t)cidx.size();
            p->max_nt = std::max(p->max_nt, nt);
            coef.resize(coef.size() + (size_t)nt * BHS_TILE_E, 0.0);
            cidx.resize(cidx.size() + (size_t)nt * BHS_TILE_E, 0);
            for (int r = 0; r < BHS_TILE_R; ++r)
                for (int c = 0; c < BHS_TILE_C; ++c) {
                    int h = tr * BHS_TILE_R + r, hp = tc * BHS_TILE_C + c;
                    if (h >= H || hp >= H) continue;
                    const std::vector<Term>& tv = terms[(size_t)h * H + hp];
                    for (size_t t = 0; t < tv.size(); ++t) {
                        size_t o = t * BHS_TILE_E + (size_t)r * BHS_TILE_C + c;
                        coef[hd.coef_off + o] = tv[t].coef;
                        cidx[hd.idx_off + o] = (uint16_t)(tv[t].idx - lo);
                    }
                }
        }
    p->coupling_bytes = (int64_t)(coef.size() * sizeof(double) + cidx.size() * sizeof(uint16_t) +
                                  p->h_tiles.size() * sizeof(bhs_tile_hdr));
    int rc;
    if ((rc = upload(&p->d_coef, coef))) return rc;
    if ((rc = upload(&p->d_cidx, cidx))) return rc;
    if ((rc = upload(&p->d_tiles, p->h_tiles))) return rc;
    return BHS_OK;
}

extern "C" int bhs_plan_create(int d, int n_end, bhs_plan_t** out) { return bhs_plan_create_tree(d, n_end, BHS_TREE_CHAIN, out); }

// tree = BHS_TREE_HOPF: the harmonic SPACE (degree < n_end on S^3) and its chain basis are those of the 'bba' plan -- the
// tree only changes where the right-hand-side quadrature samples the sphere (type-c root: n_end Gauss-Legendre nodes in
// cos 2 theta_0, 2 n_end equispaced nodes on each of the two type-a circles), which is what the reference's `caa` rows pin.
// Such a plan serves bhs_rhs_expand / bhs_plan_quadrature only (no coupling table).
extern "C" int bhs_plan_create_tree(int d, int n_end, int tree, bhs_plan_t** out) {
    if (!out) return BHS_ERR_INVALID;
    *out = nullptr;
    if (d < 2 || n_end < 1) return BHS_ERR_INVALID;
    if (tree != BHS_TREE_CHAIN && tree != BHS_TREE_HOPF) return BHS_ERR_INVALID;
    if (tree == BHS_TREE_HOPF && d != 4) return BHS_ERR_UNSUPPORTED;
    if (d - 2 > BHS_MAX_NODES) return BHS_ERR_UNSUPPORTED;  // chain types a, ba, bba, bbba, ... up to d = 8
    bhs_plan* p = new (std::nothrow) bhs_plan();
    if (!p) return BHS_ERR_ALLOC;
    p->tree = tree;
    p->d = d;
    p->s_ndim = d - 1;
    p->n_end = n_end;
    p->L2 = 2 * n_end - 1;
    p->n_bnodes = d - 2;
    build_index(p->s_ndim, n_end, p->h_idx);
    build_index(p->s_ndim, p->L2, p->h_idx2);
    p->H = (int)(p->h_idx.size() / p->s_ndim);
    p->H2 = (int)(p->h_idx2.size() / p->s_ndim);
    if (p->H2 > 65535) { delete p; return BHS_ERR_UNSUPPORTED; }
    int rc;
    std::vector<int32_t> deg(p->H), deg2(p->H2);
    for (int h = 0; h < p->H; ++h) deg[h] = std::abs(p->h_idx[(size_t)h * p->s_ndim]);
    for (int h = 0; h < p->H2; ++h) deg2[h] = std::abs(p->h_idx2[(size_t)h * p->s_ndim]);
    if ((rc = upload(&p->d_idx, p->h_idx)) || (rc = upload(&p->d_idx2, p->h_idx2)) ||
        (rc = upload(&p->d_deg, deg)) || (rc = upload(&p->d_deg2, deg2))) {
        bhs_plan_destroy(p);
        return rc;
    }
    // node recurrence tables (band L2)
    const int L2 = p->L2, nb = p->n_bnodes;
    std::vector<double> all((size_t)std::max(nb, 1) * L2, 0.0), c1((size_t)std::max(nb, 1) * L2 * L2, 0.0),
        c2((size_t)std::max(nb, 1) * L2 * L2, 0.0);
    for (int i = 0; i < nb; ++i) {
        int desc = d - 2 - i;
        for (int l = 0; l < L2; ++l) {
            all[(size_t)i * L2 + l] = (double)(1.0L / sqrtl(wallis(2 * l + desc)));
            for (int n = l + 1; n < L2; ++n) {
                ld an = sqrtl(a2coef(n, l, desc));
                ld anm1 = (n - 1 > l) ? sqrtl(a2coef(n - 1, l, desc)) : 0.0L;
                c1[((size_t)i * L2 + n) * L2 + l] = (double)(1.0L / an);
                c2[((size_t)i * L2 + n) * L2 + l] = (double)(anm1 / an);
            }
        }
    }
    if ((rc = upload(&p->d_node_all, all)) || (rc = upload(&p->d_node_c1, c1)) || (rc = upload(&p->d_node_c2, c2))) {
        bhs_plan_destroy(p);
        return rc;
    }
    if (tree == BHS_TREE_HOPF) {
        // y = (cos t0 cos t1, cos t0 sin t1, sin t0 cos t2, sin t0 sin t2), t0 in [0, pi/2]; the surface measure is
        // sin t0 cos t0 dt0 dt1 dt2 = (1/4) d(cos 2 t0) dt1 dt2
        std::vector<ld> xs, ws;
        gauss_gegenbauer(n_end, 1, xs, ws);  // Gauss-Legendre
        const int na = 2 * n_end, Q = n_end * na * na;
        p->Q = Q;
        p->h_qdirs.assign((size_t)d * Q, 0.0);
        p->h_qw.assign(Q, 0.0);
        for (int i0 = 0; i0 < n_end; ++i0) {
            const ld t0 = 0.5L * acosl(xs[i0]), c0 = cosl(t0), s0 = sinl(t0);
            for (int i1 = 0; i1 < na; ++i1) {
                const ld t1 = 2.0L * PI_L * i1 / na;
                for (int i2 = 0; i2 < na; ++i2) {
                    const ld t2 = 2.0L * PI_L * i2 / na;
                    const int q = (i0 * na + i1) * na + i2;
                    p->h_qdirs[(size_t)0 * Q + q] = (double)(c0 * cosl(t1));
                    p->h_qdirs[(size_t)1 * Q + q] = (double)(c0 * sinl(t1));
                    p->h_qdirs[(size_t)2 * Q + q] = (double)(s0 * cosl(t2));
                    p->h_qdirs[(size_t)3 * Q + q] = (double)(s0 * sinl(t2));
                    p->h_qw[q] = (double)(0.25L * ws[i0] * (PI_L / n_end) * (PI_L / n_end));
                }
            }
        }
        if ((rc = upload(&p->d_qdirs, p->h_qdirs)) || (rc = upload(&p->d_qw, p->h_qw))) {
            bhs_plan_destroy(p);
            return rc;
        }
        if (cudaMalloc((void**)&p->d_WY, (size_t)Q * p->H * sizeof(cplx)) != cudaSuccess) {
            bhs_plan_destroy(p);
            return BHS_ERR_ALLOC;
        }
    } else
    // RHS quadrature: product rule, n_end nodes per b-node, 2 n_end on the periodic node
    {
        std::vector<std::vector<ld>> ax(p->s_ndim), aw(p->s_ndim);
        for (int i = 0; i < nb; ++i) {
            std::vector<ld> xs, ws;
            gauss_gegenbauer(n_end, d - 2 - i, xs, ws);
            // descending cos <-> ascending theta is irrelevant for the sum; keep ascending x
            ax[i].resize(n_end);
            aw[i] = ws;
            for (int q = 0; q < n_end; ++q) ax[i][q] = acosl(xs[q]);
        }
        ax[p->s_ndim - 1].resize(2 * n_end);
        aw[p->s_ndim - 1].assign(2 * n_end, PI_L / n_end);
        for (int j = 0; j < 2 * n_end; ++j) ax[p->s_ndim - 1][j] = 2.0L * PI_L * j / (2 * n_end);
        int Q = 1;
        for (int i = 0; i < p->s_ndim; ++i) Q *= (int)ax[i].size();
        p->Q = Q;
        p->h_qdirs.assign((size_t)d * Q, 0.0);
        p->h_qw.assign(Q, 0.0);
        std::vector<int> cnt(p->s_ndim, 0);
        for (int q = 0; q < Q; ++q) {
            int rem = q;
            for (int i = p->s_ndim - 1; i >= 0; --i) {
                cnt[i] = rem % (int)ax[i].size();
                rem /= (int)ax[i].size();
            }
            ld w = 1.0L, prod = 1.0L;
            for (int i = 0; i < p->s_ndim; ++i) {
                w *= aw[i][cnt[i]];
                ld th = ax[i][cnt[i]];
                p->h_qdirs[(size_t)i * Q + q] = (double)(prod * cosl(th));
                prod *= sinl(th);
            }
            p->h_qdirs[(size_t)(d - 1) * Q + q] = (double)prod;
            p->h_qw[q] = (double)w;
        }
        if ((rc = upload(&p->d_qdirs, p->h_qdirs)) || (rc = upload(&p->d_qw, p->h_qw))) {
            bhs_plan_destroy(p);
            return rc;
        }
        if (cudaMalloc((void**)&p->d_WY, (size_t)Q * p->H * sizeof(cplx)) != cudaSuccess) {
            bhs_plan_destroy(p);
            return BHS_ERR_ALLOC;
        }
    }
    // 3-D fast-evaluation tables: monic recurrence q_{n+1} = x q_n - beta_{n,m} q_{n-1}, q_m = sin^m;
    // normalised P_n^m = norm_{n,m} q_n, with 1/sqrt(2 pi) folded in.  m-major order (m, n >= m).
    {
        std::vector<double> beta, norm;
        if (d == 3) {
            const int L = n_end;
            for (int m = 0; m < L; ++m) {
                ld nr = (1.0L / sqrtl(wallis(2 * m + 1))) / sqrtl(2.0L * PI_L);
                for (int n = m; n < L; ++n) {
                    ld a2 = (n > m) ? a2coef(n, m, 1) : 0.0L;
                    if (n > m) nr /= sqrtl(a2);
                    beta.push_back((double)a2);   // beta used when stepping n -> n+1 is a2coef(n, m)
                    norm.push_back((double)nr);
                }
            }
        }
        if ((rc = upload(&p->d_us_beta, beta)) || (rc = upload(&p->d_us_norm, norm))) {
            bhs_plan_destroy(p);
            return rc;
        }
    }
    if (tree == BHS_TREE_CHAIN && (rc = build_coupling(p))) {
        bhs_plan_destroy(p);
        return rc;
    }
    if ((rc = bhs_fill_WY(p)) || (rc = bhs_fill_planar_tables(p))) {
        bhs_plan_destroy(p);
        return rc;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        bhs_plan_destroy(p);
        return (int)cudaGetLastError();
    }
    *out = p;
    return BHS_OK;
}

extern "C" void bhs_plan_destroy(bhs_plan_t* p) {
    if (!p) return;
    cudaFree(p->d_idx); cudaFree(p->d_idx2); cudaFree(p->d_deg); cudaFree(p->d_deg2);
    cudaFree(p->d_node_all); cudaFree(p->d_node_c1); cudaFree(p->d_node_c2);
    cudaFree(p->d_qdirs); cudaFree(p->d_qw); cudaFree(p->d_WY);
    cudaFree(p->d_tiles); cudaFree(p->d_coef); cudaFree(p->d_cidx);
    cudaFree(p->d_us_beta); cudaFree(p->d_us_norm); cudaFree(p->d_us_rot); cudaFree(p->d_us_K);
    delete p;
}
extern "C" int bhs_plan_harm(const bhs_plan_t* p) { return p ? p->H : BHS_ERR_INVALID; }
extern "C" int bhs_plan_harm2(const bhs_plan_t* p) { return p ? p->H2 : BHS_ERR_INVALID; }
extern "C" int bhs_plan_quad_points(const bhs_plan_t* p) { return p ? p->Q : BHS_ERR_INVALID; }
extern "C" int bhs_plan_index_table(const bhs_plan_t* p, int32_t* h_out) {
    if (!p || !h_out) return BHS_ERR_INVALID;
    memcpy(h_out, p->h_idx.data(), p->h_idx.size() * sizeof(int32_t));
    return BHS_OK;
}
extern "C" int bhs_plan_quadrature(const bhs_plan_t* p, double* h_dirs, double* h_weights) {
    if (!p || !h_dirs || !h_weights) return BHS_ERR_INVALID;
    memcpy(h_dirs, p->h_qdirs.data(), p->h_qdirs.size() * sizeof(double));
    memcpy(h_weights, p->h_qw.data(), p->h_qw.size() * sizeof(double));
    return BHS_OK;
}
extern "C" int bhs_plan_coupling_stats(const bhs_plan_t* p, int64_t* nterms, int64_t* bytes) {
    if (!p) return BHS_ERR_INVALID;
    if (nterms) *nterms = p->coupling_terms;
    if (bytes) *bytes = p->coupling_bytes;
    return BHS_OK;
}
