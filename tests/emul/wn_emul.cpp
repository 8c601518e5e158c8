// wn_emul.cpp — TEST INFRASTRUCTURE ONLY. Host emulation of the device build: runs the per-thread bodies of
// lagrange_b200/csrc/wn_build_core.cuh / wn_device.cuh (the same source the sm_100a kernels inline) in sequential
// loops, so the CPU-only test tier can check the Karras topology, the arrival-counter climb, the skip-link packing and
// the folded far-field records against the oracle without a GPU. Never linked into the product.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#include "../../lagrange_b200/csrc/wn_build_core.cuh"
#include "../../lagrange_b200/csrc/wn_refbuild_core.cuh"

namespace {

struct Emul
{
    std::vector<float> v;
    std::vector<int> tri;
    std::vector<int> child, parent, arrive, ntri, size;
    std::vector<unsigned char> slot, collapsed, skip;
    std::vector<unsigned> prim, r2v, tri_order;
    std::vector<float4> local, hot, cold, tris;
    std::vector<int4> kids;
    WnBuild b;
    WnTreeView view;
    int err = 0, max_depth = 0;
};

} // namespace

// ---- K3R (WN_HIERARCHY_REFERENCE): the driver of wn_refbuild_core.cuh with a sequential backend ---------------------------
namespace {
int g_ref_sorts = 0; // order-statistic fallbacks taken by the last emul_ref_topology call (rounds that needed the sort)
struct RefHostBackend
{
    int syncs = 0;
    template <class F>
    void for_each(int64_t n, const F& f)
    {
        for (int64_t i = 0; i < n; ++i) f(i);
    }
    void bin(const WnRefState& s) { for_each(s.N, WnRefBin{s}); }
    void root_bounds(const WnRefState& s)
    {
        for (int a = 0; a < 6; ++a) s.rows[a] = a < 3 ? 0x7fffffff : (int)0x80000000;
        for (int t = 0; t < s.N; ++t) {
            float b[6];
            wn_ref_tri_box(s.tbox, (unsigned)t, b);
            for (int a = 0; a < 3; ++a) {
                s.rows[a] = std::min(s.rows[a], wn_ref_ordered(b[a]));
                s.rows[3 + a] = std::max(s.rows[3 + a], wn_ref_ordered(b[3 + a]));
            }
        }
    }
    void scan(uint32_t* d, int64_t n)
    {
        uint32_t run = 0;
        for (int64_t i = 0; i < n; ++i) {
            const uint32_t v = d[i];
            d[i] = run;
            run += v;
        }
    }
    int sort64(uint64_t* keys, unsigned* vals, unsigned*, int64_t n, int end_bit)
    {
        ++g_ref_sorts;
        const uint64_t mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1ull);
        std::vector<int64_t> order(n);
        std::iota(order.begin(), order.end(), (int64_t)0);
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t c) { return (keys[a] & mask) < (keys[c] & mask); });
        std::vector<unsigned> v2(n);
        for (int64_t i = 0; i < n; ++i) v2[i] = vals[order[i]];
        std::copy(v2.begin(), v2.end(), vals);
        return 0;
    }
    void read(int* h, const int* d, int n)
    {
        memcpy(h, d, (size_t)n * sizeof(int));
        ++syncs;
    }
    void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
    void write(void* d, const void* h, size_t bytes) { memcpy(d, h, bytes); }
};
} // namespace


extern "C" {

// child_in == nullptr: LBVH build (Morton + stable sort + Karras). Otherwise imported topology (neutral encoding).
void* emul_build(const float* v, int64_t nV, const int32_t* tri, int64_t nT, const int32_t* child_in, int64_t num_nodes, int width,
                 int leaf_size, int order, int radius_mode, int approx_single, int morton_bits, int hierarchy)
{
    // the k-d hierarchies are packed 4-ary (odd-depth internal nodes get no record); WN_EMUL_WIDE=0 (host research) keeps them binary
    const bool wide = hierarchy >= 1 && !(getenv("WN_EMUL_WIDE") && atoi(getenv("WN_EMUL_WIDE")) == 0);
    Emul* e = new Emul;
    e->v.assign(v, v + nV * 3);
    e->tri.assign(tri, tri + nT * 3);
    WnBuild& b = e->b;
    memset(&b, 0, sizeof(b));
    b.v_xyz = e->v.data();
    b.tri = e->tri.data();
    b.nV = (int)nV;
    b.nT = (int)nT;
    b.nL = (int)nT;
    b.leaf_size = leaf_size;
    b.order = order;
    b.radius_mode = radius_mode;
    b.approx_single = approx_single;
    b.err = &e->err;
    b.max_depth = &e->max_depth;
    e->prim.resize(nT);
    if (!child_in) {
        b.W = 2;
        b.nI = nT >= 2 ? (int)nT - 1 : 1;
        // K1: bounds of centroids + Morton codes (same float ops as k_centroid_bounds / k_morton)
        float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        std::vector<float> cen(nT * 3);
        for (int64_t t = 0; t < nT; ++t)
            for (int a = 0; a < 3; ++a) {
                const float c = wn_centroid_coord(v[3 * tri[3 * t] + a], v[3 * tri[3 * t + 1] + a], v[3 * tri[3 * t + 2] + a]);
                cen[3 * t + a] = c;
                if (c == c) {
                    lo[a] = std::min(lo[a], c);
                    hi[a] = std::max(hi[a], c);
                }
            }
        const int64_t nN = (int64_t)b.nI + b.nL;
        e->child.assign((size_t)b.nI * 2, -1);
        e->parent.assign(nN, -1);
        e->slot.assign(nN, 0);
        if (hierarchy == 2 && nT >= 2) {
            // K3'' (wn_kd.cuh): level-synchronous k-d build with SAH-guided split positions, explicit node ranges
            const int N = (int)nT, leaf = std::max(1, leaf_size), nI = N - 1;
            std::vector<unsigned> perm(nT), node_of(nT, 0u);
            std::iota(perm.begin(), perm.end(), 0u);
            std::vector<unsigned> start{0u, (unsigned)N};
            std::vector<int> pid{-1};
            std::vector<unsigned char> meta{0};
            e->skip.assign(b.nI, 0);
            int root_gap = 0;
            for (int level = 0;; ++level) {
                const int count = (int)pid.size();
                std::vector<float> nlo((size_t)count * 3, 3.4e38f), nhi((size_t)count * 3, -3.4e38f);
                for (int p = 0; p < N; ++p) {
                    const unsigned i = node_of[p];
                    if ((int)(start[i + 1] - start[i]) <= leaf) continue;
                    const float* c = &cen[3 * (size_t)perm[p]];
                    for (int a = 0; a < 3; ++a)
                        if (c[a] == c[a]) {
                            nlo[3 * (size_t)i + a] = std::min(nlo[3 * (size_t)i + a], c[a]);
                            nhi[3 * (size_t)i + a] = std::max(nhi[3 * (size_t)i + a], c[a]);
                        }
                }
                // RESEARCH VARIANT (host only, WN_EMUL_AXES=3): bins on all three axes, the axis with the cheapest cut wins
                static const bool kAllAxes = getenv("WN_EMUL_AXES") && atoi(getenv("WN_EMUL_AXES")) == 3;
                std::vector<int> forced_axis(count, -1), forced_left(count, 0);
                if (kAllAxes) {
                    for (int i = 0; i < count; ++i) {
                        const int s0 = (int)start[i], n = (int)(start[i + 1] - start[i]);
                        if (n <= leaf || n < WN_KDX_MIN_SAH) continue;
                        float ext3[3];
                        for (int a = 0; a < 3; ++a) ext3[a] = nhi[3 * (size_t)i + a] >= nlo[3 * (size_t)i + a] ? nhi[3 * (size_t)i + a] - nlo[3 * (size_t)i + a] : 0.0f;
                        float best_cost = 3.4e38f;
                        for (int axis = 0; axis < 3; ++axis) {
                            if (!(ext3[axis] > 0.0f)) continue;
                            float box[WN_KDX_BINS * 6];
                            int bcnt[WN_KDX_BINS];
                            for (int k = 0; k < WN_KDX_BINS; ++k) {
                                bcnt[k] = 0;
                                for (int a = 0; a < 3; ++a) box[k * 6 + a] = 3.4e38f, box[k * 6 + 3 + a] = -3.4e38f;
                            }
                            for (int j = 0; j < n; ++j) {
                                const unsigned t = perm[s0 + j];
                                const int bin = wn_kdx_bin(wn_kd_quant(cen[3 * (size_t)t + axis], nlo[3 * (size_t)i + axis], ext3[axis]));
                                ++bcnt[bin];
                                for (int a = 0; a < 3; ++a) {
                                    const float x0 = v[3 * tri[3 * t] + a], x1 = v[3 * tri[3 * t + 1] + a], x2 = v[3 * tri[3 * t + 2] + a];
                                    box[bin * 6 + a] = std::min(box[bin * 6 + a], fminf(x0, fminf(x1, x2)));
                                    box[bin * 6 + 3 + a] = std::max(box[bin * 6 + 3 + a], fmaxf(x0, fmaxf(x1, x2)));
                                }
                            }
                            const int nl2 = wn_kdx_choose(n, box, bcnt);
                            // cost of that cut (recomputed: wn_kdx_choose returns the count only)
                            float pl[3] = {3.4e38f, 3.4e38f, 3.4e38f}, ph[3] = {-3.4e38f, -3.4e38f, -3.4e38f}, sl[3] = {3.4e38f, 3.4e38f, 3.4e38f},
                                  sh[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
                            int pre = 0, k = 0;
                            for (; k < WN_KDX_BINS && pre < nl2; ++k) {
                                pre += bcnt[k];
                                for (int a = 0; a < 3; ++a) pl[a] = std::min(pl[a], box[k * 6 + a]), ph[a] = std::max(ph[a], box[k * 6 + 3 + a]);
                            }
                            if (pre != nl2) continue; // the median fallback: not a bin boundary, skip this axis
                            for (; k < WN_KDX_BINS; ++k)
                                for (int a = 0; a < 3; ++a) sl[a] = std::min(sl[a], box[k * 6 + a]), sh[a] = std::max(sh[a], box[k * 6 + 3 + a]);
                            const float cost = wn_kdx_measure(pl, ph) * nl2 + wn_kdx_measure(sl, sh) * (n - nl2);
                            if (cost < best_cost) best_cost = cost, forced_axis[i] = axis, forced_left[i] = nl2;
                        }
                    }
                }
                std::vector<uint64_t> key(nT);
                for (int p = 0; p < N; ++p) {
                    const unsigned i = node_of[p];
                    unsigned q = 0;
                    if ((int)(start[i + 1] - start[i]) > leaf) {
                        float ext3[3];
                        for (int a = 0; a < 3; ++a) ext3[a] = nhi[3 * (size_t)i + a] >= nlo[3 * (size_t)i + a] ? nhi[3 * (size_t)i + a] - nlo[3 * (size_t)i + a] : 0.0f;
                        const int axis = forced_axis[i] >= 0 ? forced_axis[i] : wn_kd_axis(ext3);
                        q = wn_kd_quant(cen[3 * (size_t)perm[p] + axis], nlo[3 * (size_t)i + axis], ext3[axis]);
                    }
                    key[p] = ((uint64_t)i << 16) | q;
                }
                {
                    std::vector<unsigned> idx2(nT);
                    std::iota(idx2.begin(), idx2.end(), 0u);
                    std::stable_sort(idx2.begin(), idx2.end(), [&](unsigned a, unsigned c2) { return key[a] < key[c2]; });
                    std::vector<unsigned> np(nT);
                    for (int64_t i = 0; i < nT; ++i) np[i] = perm[idx2[i]];
                    perm.swap(np); // node_of is unchanged: the sort only permutes inside ranges
                }
                // split decisions
                std::vector<int> nl(count, 0);
                std::vector<unsigned> off(count + 1, 0u);
                int n_split = 0;
                for (int i = 0; i < count; ++i) {
                    const int s0 = (int)start[i], n = (int)(start[i + 1] - start[i]);
                    int left = 0;
                    if (n > leaf) {
                        left = n / 2;
                        if (n >= WN_KDX_MIN_SAH) {
                            float ext3[3];
                            for (int a = 0; a < 3; ++a) ext3[a] = nhi[3 * (size_t)i + a] >= nlo[3 * (size_t)i + a] ? nhi[3 * (size_t)i + a] - nlo[3 * (size_t)i + a] : 0.0f;
                            const int axis = wn_kd_axis(ext3);
                            float box[WN_KDX_BINS * 6];
                            int bcnt[WN_KDX_BINS];
                            for (int k = 0; k < WN_KDX_BINS; ++k) {
                                bcnt[k] = 0;
                                for (int a = 0; a < 3; ++a) box[k * 6 + a] = 3.4e38f, box[k * 6 + 3 + a] = -3.4e38f;
                            }
                            for (int j = 0; j < n; ++j) {
                                const unsigned t = perm[s0 + j];
                                const int bin = wn_kdx_bin(wn_kd_quant(cen[3 * (size_t)t + axis], nlo[3 * (size_t)i + axis], ext3[axis]));
                                ++bcnt[bin];
                                for (int a = 0; a < 3; ++a) {
                                    const float x0 = v[3 * tri[3 * t] + a], x1 = v[3 * tri[3 * t + 1] + a], x2 = v[3 * tri[3 * t + 2] + a];
                                    const float lo2 = fminf(x0, fminf(x1, x2)), hi2 = fmaxf(x0, fmaxf(x1, x2));
                                    if (lo2 == lo2) box[bin * 6 + a] = std::min(box[bin * 6 + a], lo2);
                                    if (hi2 == hi2) box[bin * 6 + 3 + a] = std::max(box[bin * 6 + 3 + a], hi2);
                                }
                            }
                            left = wn_kdx_choose(n, box, bcnt);
                            if (forced_axis[i] >= 0) left = forced_left[i];
                        }
                        ++n_split;
                    }
                    nl[i] = left;
                    off[i + 1] = off[i] + (left > 0 ? 2u : 1u);
                    if (level == 0) root_gap = left > 0 ? s0 + left - 1 : N / 2 - 1;
                }
                // scatter
                const int total = (int)off[count];
                std::vector<unsigned> nstart(total + 1);
                std::vector<int> npid(total);
                std::vector<unsigned char> nmeta(total);
                for (int i = 0; i < count; ++i) {
                    const int s0 = (int)start[i], n = (int)(start[i + 1] - start[i]);
                    const int side = meta[i] & 1;
                    const bool linked = (meta[i] & 2) != 0;
                    const unsigned j = off[i];
                    const int left = nl[i];
                    if (left > 0) {
                        const int id = wn_kdx_gap_id(s0 + left - 1, root_gap);
                        if (pid[i] >= 0) {
                            e->child[2 * (size_t)pid[i] + side] = id;
                            e->parent[id] = pid[i];
                            e->slot[id] = (unsigned char)side;
                        }
                        e->skip[id] = (level & 1) ? 1 : 0;
                        const int cs[2] = {s0, s0 + left}, cm[2] = {left, n - left};
                        for (int c = 0; c < 2; ++c) {
                            nstart[j + c] = (unsigned)cs[c];
                            npid[j + c] = id;
                            if (cm[c] == 1) {
                                e->child[2 * (size_t)id + c] = nI + cs[c];
                                e->parent[nI + cs[c]] = id;
                                e->slot[nI + cs[c]] = (unsigned char)c;
                                nmeta[j + c] = (unsigned char)(c | 2);
                            } else {
                                nmeta[j + c] = (unsigned char)c;
                            }
                        }
                    } else {
                        if (!linked) wn_kdx_emit_halving(N, s0, n, pid[i], side, root_gap, e->child.data(), e->parent.data(), e->slot.data(), e->skip.data());
                        nstart[j] = (unsigned)s0;
                        npid[j] = pid[i];
                        nmeta[j] = (unsigned char)(side | 2);
                    }
                }
                nstart[total] = (unsigned)N;
                for (int p = 0; p < N; ++p) {
                    const unsigned i = node_of[p];
                    node_of[p] = off[i] + ((nl[i] > 0 && p >= (int)start[i] + nl[i]) ? 1u : 0u);
                }
                if (n_split == 0) break;
                start.swap(nstart);
                pid.swap(npid);
                meta.swap(nmeta);
            }
            e->prim = perm;
            if (wide) b.skip = e->skip.data();
        } else if (hierarchy == 1 && nT >= 2) {
            // K3' (wn_kd.cuh): per level, node centroid bounds -> (path, 16-bit coordinate) keys -> stable sort
            std::vector<unsigned> perm(nT);
            std::iota(perm.begin(), perm.end(), 0u);
            const int levels = wn_kd_levels((int)nT, leaf_size);
            for (int l = 0; l < levels; ++l) {
                const int nodes = 1 << l;
                std::vector<float> nlo((size_t)nodes * 3, 3.4e38f), nhi((size_t)nodes * 3, -3.4e38f);
                std::vector<uint64_t> key(nT);
                for (int pass = 0; pass < 2; ++pass)
                    for (int p = 0; p < (int)nT; ++p) {
                        int rlo, rn;
                        unsigned path;
                        wn_kd_locate((int)nT, p, l, rlo, rn, path);
                        const float* c = &cen[3 * (size_t)perm[p]];
                        if (pass == 0) {
                            if (rn < 2) continue;
                            for (int a = 0; a < 3; ++a)
                                if (c[a] == c[a]) {
                                    nlo[3 * (size_t)path + a] = std::min(nlo[3 * (size_t)path + a], c[a]);
                                    nhi[3 * (size_t)path + a] = std::max(nhi[3 * (size_t)path + a], c[a]);
                                }
                        } else {
                            unsigned q = 0;
                            if (rn >= 2) {
                                float ext3[3];
                                for (int a = 0; a < 3; ++a)
                                    ext3[a] = nhi[3 * (size_t)path + a] >= nlo[3 * (size_t)path + a] ? nhi[3 * (size_t)path + a] - nlo[3 * (size_t)path + a] : 0.0f;
                                const int axis = wn_kd_axis(ext3);
                                q = wn_kd_quant(c[axis], nlo[3 * (size_t)path + axis], ext3[axis]);
                            }
                            key[p] = ((uint64_t)path << 16) | q;
                        }
                    }
                std::vector<unsigned> idx2(nT);
                std::iota(idx2.begin(), idx2.end(), 0u);
                std::stable_sort(idx2.begin(), idx2.end(), [&](unsigned a, unsigned c2) { return key[a] < key[c2]; });
                std::vector<unsigned> np(nT);
                for (int64_t i = 0; i < nT; ++i) np[i] = perm[idx2[i]];
                perm.swap(np);
            }
            e->prim = perm;
            e->skip.assign(b.nI, 0);
            for (int g = 0; g < (int)nT - 1; ++g)
                wn_kd_emit_node((int)nT, g, e->child.data(), e->parent.data(), e->slot.data(), wide ? e->skip.data() : nullptr, leaf_size);
            if (wide) b.skip = e->skip.data();
        } else {
        const float ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
        const float inv = ext > 0.0f ? 1.0f / ext : 0.0f;
        std::vector<uint64_t> keys(nT);
        const int bpa = morton_bits == 63 ? 21 : 10;
        for (int64_t t = 0; t < nT; ++t)
            keys[t] = wn_morton(wn_unit_coord(cen[3 * t], lo[0], inv), wn_unit_coord(cen[3 * t + 1], lo[1], inv),
                                wn_unit_coord(cen[3 * t + 2], lo[2], inv), bpa);
        // K2: stable sort by key (what a stable LSD radix sort produces)
        std::vector<unsigned> idx(nT);
        std::iota(idx.begin(), idx.end(), 0u);
        std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned c) { return keys[a] < keys[c]; });
        std::vector<uint64_t> sorted(nT);
        for (int64_t i = 0; i < nT; ++i) sorted[i] = keys[idx[i]];
        e->prim = idx;
        if (nT >= 2) {
            for (int i = 0; i < (int)nT - 1; ++i) wn_lbvh_node(sorted.data(), (int)nT, i, e->child.data(), e->parent.data(), e->slot.data());
        } else {
            e->child[0] = 1;
            e->parent[1] = 0;
        }
        }
        b.child = e->child.data();
        b.parent = e->parent.data();
        b.slot = e->slot.data();
    } else {
        b.W = width;
        b.nI = (int)num_nodes;
        const int64_t nN = (int64_t)b.nI + b.nL;
        std::iota(e->prim.begin(), e->prim.end(), 0u);
        e->child.assign((size_t)b.nI * b.W, -1);
        e->parent.assign(nN, -1);
        e->slot.assign(nN, 0);
        b.child = e->child.data();
        b.parent = e->parent.data();
        b.slot = e->slot.data();
        std::vector<int> seen(nN, 0);
        for (int i = 0; i < b.nI; ++i) wn_import_node(b, child_in, seen.data(), i);
        for (int i = 0; i < (int)nN; ++i) wn_import_check(b, seen.data(), i);
    }
    b.prim = e->prim.data();
    const int64_t nN = (int64_t)b.nI + b.nL;
    e->local.resize((size_t)nN * 9);
    e->arrive.assign(b.nI, 0);
    e->ntri.assign(nN, 0);
    e->size.assign(nN, 0);
    e->collapsed.assign(b.nI, 0);
    e->r2v.assign(nN, 0);
    b.local = e->local.data();
    b.arrive = e->arrive.data();
    b.ntri = e->ntri.data();
    b.size = e->size.data();
    b.collapsed = e->collapsed.data();
    b.r2v = e->r2v.data();
    if (e->err == 0) {
        for (int l = 0; l < b.nL; ++l) wn_climb_leaf(b, l);
        if (radius_mode == 1)
            for (int l = 0; l < b.nL; ++l) wn_vertex_radius_leaf(b, l);
    }
    // RESEARCH VARIANT (host only, WN_EMUL_WIDE_RULE=1; used with tests/tools/hierarchy_cost.py): choose the records that are
    // dropped for the 4-ary packing like the reference builder forms its nodes (split in two, then split the
    // child with the largest area * count again until there are four) instead of by depth parity. The `kids` table written by
    // wn_pack_node is not valid for this variant (it expands one level only); the depth-first order and the skip links are.
    if (getenv("WN_EMUL_WIDE_RULE") && atoi(getenv("WN_EMUL_WIDE_RULE")) == 1 && !child_in && hierarchy >= 1 && e->err == 0 && nT >= 2) {
        e->skip.assign(b.nI, 0);
        auto weight = [&](int c) {
            const float* r = reinterpret_cast<const float*>(&e->local[(size_t)c * 9]);
            const float dx = r[3] - r[0], dy = r[4] - r[1], dz = r[5] - r[2];
            return (dx * dy + dy * dz + dz * dx) * (float)e->ntri[c];
        };
        auto expandable = [&](int c) { return c < b.nI && !e->collapsed[c]; };
        std::vector<int> queue{0};
        for (size_t qi = 0; qi < queue.size(); ++qi) {
            const int X = queue[qi];
            std::vector<int> S;
            for (int s2 = 0; s2 < b.W; ++s2)
                if (e->child[(size_t)X * b.W + s2] >= 0) S.push_back(e->child[(size_t)X * b.W + s2]);
            for (int rep = 0; rep < 2 && S.size() < 4; ++rep) {
                int bi = -1;
                float bw = -1.0f;
                for (size_t k = 0; k < S.size(); ++k)
                    if (expandable(S[k]) && weight(S[k]) > bw) bw = weight(S[k]), bi = (int)k;
                if (bi < 0) break;
                const int c = S[bi];
                e->skip[c] = 1;
                S.erase(S.begin() + bi);
                for (int s2 = 0; s2 < b.W; ++s2)
                    if (e->child[(size_t)c * b.W + s2] >= 0) S.push_back(e->child[(size_t)c * b.W + s2]);
            }
            for (int c : S)
                if (expandable(c)) queue.push_back(c);
        }
        b.skip = e->skip.data();
        // sizes again, bottom-up (reverse breadth-first order over the whole binary tree)
        std::vector<int> order{0};
        for (size_t qi = 0; qi < order.size(); ++qi) {
            const int X = order[qi];
            if (X >= b.nI) continue;
            for (int s2 = 0; s2 < b.W; ++s2)
                if (e->child[(size_t)X * b.W + s2] >= 0) order.push_back(e->child[(size_t)X * b.W + s2]);
        }
        for (size_t qi = order.size(); qi-- > 0;) {
            const int X = order[qi];
            if (X >= b.nI || e->collapsed[X]) {
                e->size[X] = 1;
                continue;
            }
            int sz = e->skip[X] ? 0 : 1;
            for (int s2 = 0; s2 < b.W; ++s2)
                if (e->child[(size_t)X * b.W + s2] >= 0) sz += e->size[e->child[(size_t)X * b.W + s2]];
            e->size[X] = sz;
        }
    }
    const int n_entries = (e->err == 0 && nT > 0) ? e->size[0] : 0;
    e->hot.resize((size_t)n_entries * 2);
    e->cold.resize((size_t)n_entries * 4);
    b.hot = e->hot.data();
    b.cold = e->cold.data();
    e->kids.resize(n_entries);
    b.kids = e->kids.data();
    e->tris.resize((size_t)nT * 3);
    e->tri_order.resize(nT);
    b.tris = e->tris.data();
    b.tri_order = e->tri_order.data();
    if (e->err == 0 && nT > 0 && e->ntri[0] == b.nL)
        for (int node = 0; node < (int)nN; ++node) wn_pack_node(b, node);
    else if (nT > 0 && e->err == 0)
        e->err = WN_ERR_TOPOLOGY_BAD_CHILD;
    e->view.hot = e->hot.data();
    e->view.cold = e->cold.data();
    e->view.kids = e->kids.data();
    e->view.tri = e->tris.data();
    e->view.n_entries = n_entries;
    e->view.n_tris = (int)nT;
    return e;
}

// out_child: at least max(1, nT) * 4 ints. Returns the number of nodes (0 for an empty mesh).
int64_t emul_ref_topology(const float* v, int64_t nV, const int32_t* tri, int64_t nT, int32_t* out_child, int32_t* out_levels, int32_t* out_syncs)
{
    (void)nV;
    g_ref_sorts = 0;
    if (nT <= 0) return 0;
    const WnRefLayout L = wn_ref_layout(nT, 64);
    std::vector<char> mem(L.total + 256);
    char* base = mem.data() + (256 - ((uintptr_t)mem.data() & 255)) % 256;
    std::vector<int> child_out(((size_t)nT + 2) * 4, -1);
    WnRefState rs;
    memset(&rs, 0, sizeof(rs));
    rs.N = (int)nT;
    rs.tbox = (const float4*)(base + L.tbox);
    rs.idx = (unsigned*)(base + L.idx);
    rs.idx_alt = (unsigned*)(base + L.idx_alt);
    rs.owner = (int*)(base + L.owner);
    rs.flag = (uint32_t*)(base + L.flag);
    rs.nodes = (WnRefNode*)(base + L.nodes);
    rs.next = (WnRefNode*)(base + L.next);
    rs.tasks = (WnRefTask*)(base + L.tasks);
    rs.rows = (int*)(base + L.rows);
    rs.copen = (uint32_t*)(base + L.copen);
    rs.cnode = (uint32_t*)(base + L.cnode);
    rs.child_tmp = (int*)(base + L.child_tmp);
    rs.info_start = (int*)(base + L.info_start);
    rs.info_depth = (int*)(base + L.info_depth);
    rs.info_chain = (int*)(base + L.info_chain);
    rs.cnt_start = (uint32_t*)(base + L.cnt_start);
    rs.final_of = (int*)(base + L.final_of);
    rs.keys = (uint64_t*)(base + L.keys);
    rs.res = (int*)(base + L.res);
    rs.child_out = child_out.data();
    RefHostBackend B;
    B.for_each(nT, WnRefTriBoxes{v, tri, (float4*)(base + L.tbox), rs.idx, rs.owner});
    int n_nodes = 0, n_levels = 0;
    if (!wn_ref_build_topology(B, rs, &n_nodes, &n_levels)) return -1;
    memcpy(out_child, child_out.data(), (size_t)n_nodes * 4 * sizeof(int));
    if (out_levels) *out_levels = n_levels;
    if (out_syncs) *out_syncs = B.syncs;
    return n_nodes;
}

int emul_ref_last_sorts(void)
{
    return g_ref_sorts;
}

void emul_destroy(void* h)
{
    delete static_cast<Emul*>(h);
}
int emul_error(void* h)
{
    return static_cast<Emul*>(h)->err;
}
int64_t emul_num_internal(void* h)
{
    return static_cast<Emul*>(h)->b.nI;
}
int64_t emul_num_entries(void* h)
{
    return static_cast<Emul*>(h)->view.n_entries;
}
int emul_max_depth(void* h)
{
    return static_cast<Emul*>(h)->max_depth;
}
int emul_width(void* h)
{
    return static_cast<Emul*>(h)->b.W;
}

// neutral encoding, [nI * W]
void emul_get_topology(void* h, int32_t* out)
{
    Emul* e = static_cast<Emul*>(h);
    for (size_t k = 0; k < e->child.size(); ++k) {
        const int c = e->child[k];
        out[k] = c < 0 ? WN_CHILD_EMPTY : (c >= e->b.nI ? wn_enc_tri((int)e->prim[c - e->b.nI]) : c);
    }
}

void emul_get_ref23(void* h, int64_t first, int64_t count, float* out)
{
    Emul* e = static_cast<Emul*>(h);
    for (int64_t k = 0; k < count; ++k) {
        WnLocal d;
        wn_load_local(e->local.data() + (first + k) * 9, d, false);
        wn_local_to_ref23(d, out + k * 23);
    }
}

// packed arrays: rec [6][n_entries][4], link [n_entries], tris [nT][3][4], tri_order [nT]
void emul_get_packed(void* h, float* rec, int32_t* link, float* tris, uint32_t* tri_order)
{
    Emul* e = static_cast<Emul*>(h);
    // de-interleave the hot/cold layout back into the six logical float4 arrays + link
    const size_t n = e->kids.size();
    for (size_t i = 0; i < n; ++i) {
        memcpy(rec + (0 * n + i) * 4, &e->hot[2 * i], sizeof(float4));
        memcpy(rec + (1 * n + i) * 4, &e->hot[2 * i + 1], sizeof(float4));
        for (int k = 0; k < 4; ++k) memcpy(rec + ((2 + k) * n + i) * 4, &e->cold[4 * i + k], sizeof(float4));
        link[i] = wn_float_as_int(e->hot[2 * i + 1].w);
    }
    memcpy(tris, e->tris.data(), e->tris.size() * sizeof(float4));
    memcpy(tri_order, e->tri_order.data(), e->tri_order.size() * sizeof(unsigned));
}

void emul_get_kids(void* h, int32_t* out)
{
    Emul* e = static_cast<Emul*>(h);
    memcpy(out, e->kids.data(), e->kids.size() * sizeof(int4));
}

void emul_query(void* h, const float* q, int64_t n, float beta, float* out, uint64_t* counters)
{
    Emul* e = static_cast<Emul*>(h);
    unsigned long long cnt[3] = {0, 0, 0};
    for (int64_t i = 0; i < n; ++i)
        out[i] = wn_traverse_point(e->view, q[3 * i], q[3 * i + 1], q[3 * i + 2], beta * beta, counters ? cnt : nullptr);
    if (counters) {
        counters[0] = cnt[0];
        counters[1] = cnt[1];
        counters[2] = cnt[2];
    }
}

// Cost model of the tiled query path (k_tile_plan + k_tile_query) on this packed tree, evaluated on the host for every
// `tile_stride`-th 8x8x8 tile of a lattice: the same classification rules as the plan kernel (tile bounding sphere, far set /
// direct / dropped / conditional) and the same warp-level walk as warp_traverse<2, ., true> (a 4x4x4 sub-block per warp, two
// groups of 32 points, per-point resume index, the warp visits an item when any point needs it). Counts what determines the
// kernels' instruction counts; it lets hierarchy variants be compared without a GPU (tests/tools/hierarchy_cost.py).
// out[0] tiles, [1] conditional walk steps (per warp, summed), [2] evaluations executed (per 32-point group), [3] far-set
// records, [4] direct records, [5] conditional items (plan list length), [6] exact triangle evaluations (per group), [7] warps
void emul_tile_cost(void* h, const float* origin, const float* spacing, const int64_t* dims, int tile_stride, float beta, float kappa,
                    const int* warp_shape /* points per warp along x, y, z ({4,4,4} is the kernel's) */,
                    const int* tile_shape /* points per tile along x, y, z ({8,8,8} is the kernel's); multiples of warp_shape */, double* out)
{
    Emul* e = static_cast<Emul*>(h);
    const WnTreeView& t = e->view;
    const int n = t.n_entries;
    const float beta2 = beta * beta;
    for (int k = 0; k < 8; ++k) out[k] = 0.0;
    if (n == 0) return;
    const int TX = tile_shape[0], TY = tile_shape[1], TZ = tile_shape[2];
    const int tx = (int)((dims[0] + TX - 1) / TX), ty = (int)((dims[1] + TY - 1) / TY), tz = (int)((dims[2] + TZ - 1) / TZ);
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel for collapse(2) schedule(dynamic, 4) reduction(+ : acc[:8])
    for (int bz = 0; bz < tz; bz += tile_stride)
        for (int by = 0; by < ty; by += tile_stride)
            for (int bx = 0; bx < tx; bx += tile_stride) {
                // tile sphere, as in k_tile_plan
                float lo[3], hi[3];
                const int b3[3] = {bx, by, bz}, T3[3] = {TX, TY, TZ};
                for (int a = 0; a < 3; ++a) {
                    const float l = wn_lattice_coord(origin[a], spacing[a], b3[a] * T3[a]), u = wn_lattice_coord(origin[a], spacing[a], b3[a] * T3[a] + T3[a] - 1);
                    lo[a] = std::min(l, u);
                    hi[a] = std::max(l, u);
                }
                float c[3], r2 = 0.0f;
                for (int a = 0; a < 3; ++a) {
                    c[a] = 0.5f * (lo[a] + hi[a]);
                    const float hh = 0.5f * (hi[a] - lo[a]);
                    r2 += hh * hh;
                }
                const float ra = sqrtf(r2) * 1.0001f + 1e-30f;
                acc[0] += 1;
                const int wx = warp_shape[0], wy = warp_shape[1], wz = warp_shape[2];
                const int nsx = TX / wx, nsy = TY / wy, NP = wx * wy * wz, nsub = (TX * TY * TZ) / NP; // NP = 32 * (queries per lane)
                for (int sub = 0; sub < nsub; ++sub) {
                    float qx[256], qy[256], qz[256];
                    int skip[256];
                    const int sx = sub % nsx, sy = (sub / nsx) % nsy, sz = sub / (nsx * nsy);
                    for (int p = 0; p < NP; ++p) {
                        // x fastest, z slowest: the first 32 points (group k = 0) are the lower half along the slowest dimension
                        const int lx = p % wx, ly = (p / wx) % wy, lz = p / (wx * wy);
                        qx[p] = wn_lattice_coord(origin[0], spacing[0], bx * TX + sx * wx + lx);
                        qy[p] = wn_lattice_coord(origin[1], spacing[1], by * TY + sy * wy + ly);
                        qz[p] = wn_lattice_coord(origin[2], spacing[2], bz * TZ + sz * wz + lz);
                        skip[p] = 0;
                    }
                    acc[7] += 1;
                    int i = n > 1 ? 1 : 0, mixed_until = 0;
                    while (i < n) {
                        const float4 f0 = t.hot[2 * (int64_t)i], f1 = t.hot[2 * (int64_t)i + 1];
                        const bool leaf = wn_float_as_int(f0.w) < 0;
                        const int lk = wn_float_as_int(f1.w);
                        const int after = leaf ? i + 1 : lk;
                        const float thr = fabsf(f0.w) * beta2;
                        const float dx = c[0] - f0.x, dy = c[1] - f0.y, dz = c[2] - f0.z;
                        const float D = sqrtf(dx * dx + dy * dy + dz * dz);
                        const float dm = D - ra, dp = D + ra;
                        const bool allfar = dm > 0.0f && dm * dm > thr * 1.0001f;
                        const bool allnear = dp * dp <= thr * 0.9999f;
                        const bool in_mixed = i < mixed_until;
                        if (!in_mixed && allfar) {
                            if (sub == 0) {
                                if (D >= kappa * ra && D - sqrtf(fabsf(f0.w)) >= 0.5f * kappa * ra)
                                    acc[3] += 1;
                                else
                                    acc[4] += 1;
                            }
                            i = after;
                            continue;
                        }
                        if (allnear && !leaf) { // dropped: children expanded
                            i = i + 1;
                            continue;
                        }
                        if (!in_mixed && allnear && leaf) { // exact for everybody, no test
                            const int count = (lk & (WN_MAX_LEAF_SIZE - 1)) + 1;
                            acc[6] += (NP / 32.0) * count;
                            i = i + 1;
                            continue;
                        }
                        if (!in_mixed && !leaf) mixed_until = lk;
                        if (sub == 0) acc[5] += 1; // (an under-count for items only other sub-blocks reach; fine for a model)
                        bool any_active = false, anyfar[8] = {false}, anynear_k[8] = {false};
                        for (int p = 0; p < NP; ++p) {
                            if (i < skip[p]) continue;
                            any_active = true;
                            const float rx = qx[p] - f0.x, ry = qy[p] - f0.y, rz = qz[p] - f0.z;
                            const float l2 = rx * rx + ry * ry + rz * rz;
                            const bool nr = allfar ? false : (l2 <= thr);
                            const int k = p >> 5;
                            if (nr) {
                                anynear_k[k] = true;
                            } else {
                                anyfar[k] = true;
                                skip[p] = after;
                            }
                        }
                        // the real warp only visits the item if some lane is active: when none is, it jumped past it earlier
                        if (!any_active) {
                            i = after;
                            continue;
                        }
                        acc[1] += 1;
                        bool anynear = false;
                        int near_groups = 0;
                        for (int k = 0; k < NP / 32; ++k) {
                            acc[2] += anyfar[k] ? 1 : 0;
                            anynear = anynear || anynear_k[k];
                            near_groups += anynear_k[k] ? 1 : 0;
                        }
                        if (leaf) {
                            if (anynear) acc[6] += near_groups * ((lk & (WN_MAX_LEAF_SIZE - 1)) + 1);
                            i = i + 1;
                        } else {
                            i = anynear ? i + 1 : after;
                        }
                    }
                }
            }
    for (int k = 0; k < 8; ++k) out[k] = acc[k];
}

int emul_inside_from_omega(float omega)
{
    return wn_inside_from_omega(omega) ? 1 : 0;
}

// distance from each point to each triangle (a, b, c rows of 3 floats), squared, with the device's float routine
void emul_point_tri_dist2(const float* p, const float* tri9, int64_t n, float* out)
{
    for (int64_t i = 0; i < n; ++i) {
        const float* t = tri9 + 9 * i;
        out[i] = wn_point_tri_dist2(p[3 * i], p[3 * i + 1], p[3 * i + 2], make_float4(t[0], t[1], t[2], 0.0f), make_float4(t[3], t[4], t[5], 0.0f),
                                    make_float4(t[6], t[7], t[8], 0.0f));
    }
}

float emul_lattice_coord(float origin, float spacing, int i)
{
    return wn_lattice_coord(origin, spacing, i);
}

} // extern "C"
