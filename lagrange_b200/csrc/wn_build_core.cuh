// wn_build_core.cuh — per-thread bodies of the hierarchy build (K1, K3, K4, K5) as host/device functions.
//
// Every body is "what thread `tid` does"; the sm_100a kernels in wn_build.cuh call them with tid = global thread id,
// and tests/emul/wn_emul.cpp calls the same bodies in a sequential loop on the host (test infrastructure: it lets the
// CPU-only test tier check Karras topology, the atomic-counter climb, the skip-link packing and the record folding
// against the oracle without a GPU). The device path never runs on the host.
//
// Node ids: internal nodes 0 .. nI-1 (root = 0), leaf l (one triangle) = nI + l. `prim[l]` = triangle id of leaf l.
#pragma once

#include "wn_device.cuh"

#if defined(__CUDA_ARCH__)
#define WN_ATOMIC_ADD_INT(p, v) atomicAdd((p), (v))
#define WN_ATOMIC_MAX_INT(p, v) atomicMax((p), (v))
#define WN_ATOMIC_MAX_U32(p, v) atomicMax((p), (v))
#define WN_THREADFENCE() __threadfence()
#define WN_LDCG_INT(p) __ldcg((p))
#define WN_LDCG_F4(p) __ldcg((p))
#else
static inline int wn_host_atomic_add(int* p, int v)
{
    const int o = *p;
    *p = o + v;
    return o;
}
static inline int wn_host_atomic_max(int* p, int v)
{
    const int o = *p;
    if (v > o) *p = v;
    return o;
}
static inline unsigned wn_host_atomic_max_u(unsigned* p, unsigned v)
{
    const unsigned o = *p;
    if (v > o) *p = v;
    return o;
}
#define WN_ATOMIC_ADD_INT(p, v) wn_host_atomic_add((p), (v))
#define WN_ATOMIC_MAX_INT(p, v) wn_host_atomic_max((p), (v))
#define WN_ATOMIC_MAX_U32(p, v) wn_host_atomic_max_u((p), (v))
#define WN_THREADFENCE() ((void)0)
#define WN_LDCG_INT(p) (*(p))
#define WN_LDCG_F4(p) (*(p))
#endif

#define WN_MAX_WIDTH 4
#define WN_ERR_TOPOLOGY_BAD_CHILD 1
#define WN_ERR_TOPOLOGY_DEPTH 2

// ---- k-d hierarchy with SAH-guided split positions (K3'', wn_kd.cuh): explicit node ranges, level by level ---------------
// A node is a range [start, start + n) of the current triangle order. Internal nodes are numbered by the gap they split
// (start + nl - 1), with the root's gap and gap 0 swapped so that the root is node 0.
#ifndef WN_KDX_MIN_SAH
#define WN_KDX_MIN_SAH 8 /* ranges shorter than this are split at the median (measured on cfg2 with the count-eighths rule: 8 -> 6.25, 16 -> 6.20, 64 -> 6.17 G q/s) */
#endif

WN_HD int wn_kdx_gap_id(int gap, int root_gap)
{
    return gap == root_gap ? 0 : (gap == 0 ? root_gap : gap);
}

WN_HD float wn_kdx_half_area(const float* lo, const float* hi)
{
    if (!(hi[0] >= lo[0]) || !(hi[1] >= lo[1]) || !(hi[2] >= lo[2])) return 0.0f; // empty (or NaN) box
    const float dx = WN_SUB(hi[0], lo[0]), dy = WN_SUB(hi[1], lo[1]), dz = WN_SUB(hi[2], lo[2]);
    return WN_ADD(WN_ADD(WN_MUL(dx, dy), WN_MUL(dy, dz)), WN_MUL(dz, dx));
}

// How likely a node with this box is opened by a query. WN_KDX_COST = 0: half the surface area (the ray-tracing SAH);
// 1: diagonal cubed, 2: diagonal squared (the acceptance sphere of a record has radius beta * R with R ~ half the diagonal, and
// point queries that fill a volume open the node in proportion to that sphere's volume).
#ifndef WN_KDX_COST
#define WN_KDX_COST 0
#endif
WN_HD float wn_kdx_measure(const float* lo, const float* hi)
{
#if WN_KDX_COST == 0
    return wn_kdx_half_area(lo, hi);
#else
    if (!(hi[0] >= lo[0]) || !(hi[1] >= lo[1]) || !(hi[2] >= lo[2])) return 0.0f;
    const float dx = WN_SUB(hi[0], lo[0]), dy = WN_SUB(hi[1], lo[1]), dz = WN_SUB(hi[2], lo[2]);
    const float d2 = WN_ADD(WN_ADD(WN_MUL(dx, dx), WN_MUL(dy, dy)), WN_MUL(dz, dz));
#if WN_KDX_COST == 1
    return WN_MUL(d2, WN_SQRT(d2));
#elif WN_KDX_COST == 3
    return WN_MUL(d2, d2);
#elif WN_KDX_COST == 4
    { const float ar = WN_ADD(WN_ADD(WN_MUL(dx, dy), WN_MUL(dy, dz)), WN_MUL(dz, dx)); return WN_MUL(ar, WN_SQRT(ar)); }
#else
    return d2;
#endif
#endif
}

// Bins of equal WIDTH along the split axis of a range (the reference builder's rule, SURVEY.md A.6: "16 bins along the longest
// axis of the centre bounds"); q16 is the 16-bit position of a centroid inside the range's centroid bounds (wn_kd_quant).
#define WN_KDX_BINS 16
#define WN_KDX_ROW (WN_KDX_BINS * 7) /* ints per range in the bin table: BINS boxes (min xyz, max xyz) + BINS counts */
WN_HD int wn_kdx_bin(unsigned q16)
{
    const int b = (int)((q16 * (unsigned)WN_KDX_BINS) >> 16);
    return b < WN_KDX_BINS ? b : WN_KDX_BINS - 1;
}

// Left count of a range of n >= WN_KDX_MIN_SAH triangles sorted along its axis, from the triangle boxes (box[k*6 + 0..2] = min,
// + 3..5 = max) and centroid counts of its bins: the cut at a bin boundary with the least
//   measure(left) * n_left + measure(right) * n_right,
// both sides holding at least n/8 triangles (bounds the depth like the eighths did); the first such cut on ties, the median
// if no boundary qualifies (all centroids in one bin).
WN_HD int wn_kdx_choose(int n, const float* box, const int* cnt)
{
    float slo[WN_KDX_BINS][3], shi[WN_KDX_BINS][3]; // boxes of bins k..BINS-1
    for (int k = WN_KDX_BINS - 1; k >= 0; --k)
        for (int a = 0; a < 3; ++a) {
            const float l = box[k * 6 + a], h = box[k * 6 + 3 + a];
            slo[k][a] = k == WN_KDX_BINS - 1 ? l : wn_min(slo[k + 1][a], l);
            shi[k][a] = k == WN_KDX_BINS - 1 ? h : wn_max(shi[k + 1][a], h);
        }
    float pl[3] = {3.4e38f, 3.4e38f, 3.4e38f}, ph[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    const int gmin = n / 8 > 1 ? n / 8 : 1;
    int best_nl = n / 2, pre = 0;
    float best = 3.4e38f;
    for (int k = 1; k < WN_KDX_BINS; ++k) {
        for (int a = 0; a < 3; ++a) {
            pl[a] = wn_min(pl[a], box[(k - 1) * 6 + a]);
            ph[a] = wn_max(ph[a], box[(k - 1) * 6 + 3 + a]);
        }
        pre += cnt[k - 1];
        if (pre < gmin || n - pre < gmin) continue;
        const float cost = WN_ADD(WN_MUL(wn_kdx_measure(pl, ph), (float)pre), WN_MUL(wn_kdx_measure(slo[k], shi[k]), (float)(n - pre)));
        if (cost < best) {
            best = cost;
            best_nl = pre;
        }
    }
    return best_nl;
}

// A finished range (n <= leaf_size, or a single triangle) hangs off (pid, side): its inner structure is the implicit halving
// tree (it is collapsed into one leaf record later; only the arrays have to be complete).
WN_HD void wn_kdx_emit_halving(int N, int s, int m, int pid, int side, int root_gap, int* child, int* parent, unsigned char* slot,
                               unsigned char* skip)
{
    const int nI = N - 1;
    int st_s[40], st_m[40], st_p[40], st_side[40], top = 0;
    st_s[0] = s, st_m[0] = m, st_p[0] = pid, st_side[0] = side, top = 1;
    while (top > 0) {
        --top;
        const int cs = st_s[top], cm = st_m[top], cp = st_p[top], cside = st_side[top];
        const int id = cm == 1 ? nI + cs : wn_kdx_gap_id(cs + cm / 2 - 1, root_gap);
        if (cp >= 0) {
            child[2 * (size_t)cp + cside] = id;
            parent[id] = cp;
            slot[id] = (unsigned char)cside;
        }
        if (cm >= 2) {
            if (skip) skip[id] = 0;
            if (top + 2 > 40) return; // cannot happen for m <= 2^19
            st_s[top] = cs, st_m[top] = cm / 2, st_p[top] = id, st_side[top] = 0, ++top;
            st_s[top] = cs + cm / 2, st_m[top] = cm - cm / 2, st_p[top] = id, st_side[top] = 1, ++top;
        }
    }
}

// ---- balanced k-d hierarchy (K3', wn_kd.cuh): the implicit tree over the final triangle order ------------------------
// Range [lo, lo + n) and path bits of the level-`level` node that contains position p in the implicit balanced tree over N.
WN_HD void wn_kd_locate(int N, int p, int level, int& lo, int& n, unsigned& path)
{
    lo = 0;
    n = N;
    path = 0;
    for (int l = 0; l < level; ++l) {
        const int nl = n >= 2 ? n / 2 : n; // a single triangle stays where it is (path bit 0)
        if (p < lo + nl) {
            n = nl;
            path = path << 1;
        } else {
            lo += nl;
            n -= nl;
            path = (path << 1) | 1u;
        }
    }
}

// Internal node of a range [lo, lo + n), n >= 2: the index of the gap it splits (between lo + n/2 - 1 and lo + n/2). Every
// gap 0..N-2 is split by exactly one range, so gaps number the N-1 internal nodes; the root's gap and gap 0 swap so that
// the root is node 0 like in the Karras layout.
WN_HD int wn_kd_node_id(int N, int lo, int n)
{
    const int gap = lo + n / 2 - 1;
    const int root_gap = N / 2 - 1;
    return gap == root_gap ? 0 : (gap == 0 ? root_gap : gap);
}


// Number of levels whose order has to be computed: ranges of at most leaf_size triangles are collapsed into one leaf later,
// and the order inside a leaf does not matter.
WN_HD int wn_kd_levels(int N, int leaf_size)
{
    int levels = 0;
    while ((((long long)N + (1ll << levels) - 1) >> levels) > (leaf_size > 1 ? leaf_size : 1)) ++levels;
    return levels;
}

WN_HD int wn_kd_axis(const float ext[3])
{
    return (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
}

// 16-bit position of coordinate c inside [blo, blo + ext] (unfused, so host emulation and device agree)
WN_HD unsigned wn_kd_quant(float c, float blo, float ext)
{
    if (!(ext > 0.0f) || !(c == c)) return 0u;
    const float u = WN_MUL(WN_DIV(WN_SUB(c, blo), ext), 65535.0f);
    return (unsigned)(u < 0.0f ? 0.0f : (u > 65535.0f ? 65535.0f : u));
}

// Internal node that splits gap g: find its range by descending from the root, emit children / parents / slots in the
// layout of wn_lbvh_node (internal nodes 0..N-2, leaf of sorted position p = node N-1 + p). parent[] was preset to -1.
WN_HD void wn_kd_emit_node(int N, int g, int* child, int* parent, unsigned char* slot, unsigned char* skip, int leaf_size)
{
    int lo = 0, n = N, depth = 0;
    while (true) {
        const int nl = n / 2;
        const int gap = lo + nl - 1;
        if (g == gap) break;
        if (g < gap) {
            n = nl;
        } else {
            lo += nl;
            n -= nl;
        }
        ++depth;
    }
    const int id = wn_kd_node_id(N, lo, n);
    // 4-ary tree: internal nodes at odd depth get no record (unless they end up as a collapsed leaf, n <= leaf_size)
    if (skip) skip[id] = ((depth & 1) && n > (leaf_size > 1 ? leaf_size : 1)) ? 1 : 0;
    const int nI = N - 1;
    const int nl = n / 2, nr = n - nl;
    const int cl = nl >= 2 ? wn_kd_node_id(N, lo, nl) : nI + lo;
    const int cr = nr >= 2 ? wn_kd_node_id(N, lo + nl, nr) : nI + lo + nl;
    child[2 * (size_t)id] = cl;
    child[2 * (size_t)id + 1] = cr;
    parent[cl] = id;
    parent[cr] = id;
    slot[cl] = 0;
    slot[cr] = 1;
}

struct WnBuild
{
    // mesh
    const float* v_xyz;  // [nV*3]
    const int* tri;      // [nT*3]
    int nV, nT;
    // topology
    int nI, nL, W;
    int* child;          // [nI*W] node ids, -1 empty
    int* parent;         // [nI+nL], root -1
    unsigned char* slot; // [nI+nL]
    const unsigned* prim; // [nL]
    // bottom-up state
    float4* local;       // [(nI+nL) * 9] WnLocal records
    int* arrive;         // [nI] zero-initialised
    int* ntri;           // [nI+nL]
    int* size;           // [nI+nL] entries in the packed subtree (1 for leaves / collapsed nodes)
    unsigned char* collapsed; // [nI]
    const unsigned char* skip; // [nI] or null: internal nodes that get no record: their children hang off their parent (4-ary tree)
    unsigned* r2v;       // [nI+nL] vertex-radius^2 as ordered uint (float bits), zero-initialised (WN_RADIUS_VERTEX)
    int* err;            // [1] error flag
    int* max_depth;      // [1]
    // options
    int leaf_size, order, radius_mode, approx_single;
    // packed output
    float4* hot;         // [n_entries * 2]  (P, R2 | leaf) , (N, link bits)
    float4* cold;        // [n_entries * 4]  quadratic + cubic form
    int4* kids;          // [n_entries] child entry indices of internal entries
    float4* tris;        // [nT*3]
    unsigned* tri_order; // [nT] triangle id at each depth-first position
};

WN_HD void wn_store_local(float4* dst, const WnLocal& d)
{
    const float* f = (const float*)&d;
#pragma unroll
    for (int k = 0; k < 9; ++k) dst[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
}
WN_HD void wn_load_local(const float4* src, WnLocal& d, bool coherent)
{
    float* f = (float*)&d;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 v = coherent ? WN_LDCG_F4(src + k) : src[k];
        f[4 * k] = v.x;
        f[4 * k + 1] = v.y;
        f[4 * k + 2] = v.z;
        f[4 * k + 3] = v.w;
    }
}

WN_HD WnV3 wn_vertex(const float* v_xyz, int i)
{
    return wn_v3(v_xyz[3 * (int64_t)i], v_xyz[3 * (int64_t)i + 1], v_xyz[3 * (int64_t)i + 2]);
}

// ---- K3' : import a caller-supplied topology (neutral encoding -> node ids, parents, slots) ----------------------
// child_in: [nI*W], c >= 0 internal, -1 empty, <= -2 triangle -(c+2).  seen: [nI+nL] zero-initialised reference counts.
WN_HD void wn_import_node(const WnBuild& b, const int* child_in, int* seen, int i)
{
    bool empty_seen = false;
    for (int s = 0; s < b.W; ++s) {
        const int c = child_in[(int64_t)i * b.W + s];
        int node = -1;
        if (c == WN_CHILD_EMPTY) {
            empty_seen = true;
        } else if (c >= 0) {
            if (c >= b.nI || c == 0 || empty_seen)
                *b.err = WN_ERR_TOPOLOGY_BAD_CHILD;
            else
                node = c;
        } else {
            const int t = wn_dec_tri(c);
            if (t < 0 || t >= b.nL || empty_seen)
                *b.err = WN_ERR_TOPOLOGY_BAD_CHILD;
            else
                node = b.nI + t;
        }
        b.child[(int64_t)i * b.W + s] = node;
        if (node >= 0) {
            b.parent[node] = i;
            b.slot[node] = (unsigned char)s;
            WN_ATOMIC_ADD_INT(&seen[node], 1);
        }
    }
    if (b.child[(int64_t)i * b.W] < 0) *b.err = WN_ERR_TOPOLOGY_BAD_CHILD; // an internal node needs a child
}
// every non-root node must be referenced exactly once
WN_HD void wn_import_check(const WnBuild& b, const int* seen, int node)
{
    const int want = node == 0 ? 0 : 1;
    if (seen[node] != want) *b.err = WN_ERR_TOPOLOGY_BAD_CHILD;
}

// ---- K4 : per-leaf moments, then climb with arrival counters -----------------------------------------------------
WN_HD void wn_climb_leaf(const WnBuild& b, int l)
{
    const int t = (int)b.prim[l];
    const int i0 = b.tri[3 * (int64_t)t], i1 = b.tri[3 * (int64_t)t + 1], i2 = b.tri[3 * (int64_t)t + 2];
    WnLocal d;
    wn_tri_local(wn_vertex(b.v_xyz, i0), wn_vertex(b.v_xyz, i1), wn_vertex(b.v_xyz, i2), d);
    int node = b.nI + l;
    wn_store_local(b.local + (int64_t)node * 9, d);
    b.ntri[node] = 1;
    b.size[node] = 1;
    int guard = 0;
    while (true) {
        const int p = b.parent[node];
        if (p < 0) break;
        int nch = 0;
        for (int s = 0; s < b.W; ++s) nch += b.child[(int64_t)p * b.W + s] >= 0 ? 1 : 0;
        WN_THREADFENCE();
        const int old = WN_ATOMIC_ADD_INT(&b.arrive[p], 1);
        if (old + 1 < nch) break;
        // last arriver merges (children in slot order => deterministic)
        WN_THREADFENCE();
        WnLocal ch[WN_MAX_WIDTH];
        int n = 0, nt = 0, sz = (b.skip && b.skip[p]) ? 0 : 1;
        for (int s = 0; s < b.W; ++s) {
            const int c = b.child[(int64_t)p * b.W + s];
            if (c < 0) continue;
            wn_load_local(b.local + (int64_t)c * 9, ch[n], true);
            nt += WN_LDCG_INT(&b.ntri[c]);
            sz += WN_LDCG_INT(&b.size[c]);
            ++n;
        }
        WnLocal m;
        wn_merge_children(ch, n, m);
        wn_store_local(b.local + (int64_t)p * 9, m);
        const bool col = b.leaf_size > 1 && nt <= b.leaf_size;
        b.collapsed[p] = col ? 1 : 0;
        b.ntri[p] = nt;
        b.size[p] = col ? 1 : sz;
        node = p;
        if (++guard > (1 << 20)) {
            *b.err = WN_ERR_TOPOLOGY_DEPTH;
            break;
        }
    }
}

// ---- K4b : exact vertex radius (WN_RADIUS_VERTEX): every leaf pushes its farthest vertex distance to all ancestors --
WN_HD void wn_vertex_radius_leaf(const WnBuild& b, int l)
{
    const int t = (int)b.prim[l];
    const WnV3 a = wn_vertex(b.v_xyz, b.tri[3 * (int64_t)t]);
    const WnV3 bb = wn_vertex(b.v_xyz, b.tri[3 * (int64_t)t + 1]);
    const WnV3 c = wn_vertex(b.v_xyz, b.tri[3 * (int64_t)t + 2]);
    int node = b.nI + l;
    int guard = 0;
    while (node >= 0) {
        // P lives at floats 6,7,8 of the WnLocal record
        const float* rec = (const float*)(b.local + (int64_t)node * 9);
        const float px = rec[6], py = rec[7], pz = rec[8];
        float r2 = 0.0f;
        {
            const float dx = a.x - px, dy = a.y - py, dz = a.z - pz;
            r2 = wn_max(r2, dx * dx + dy * dy + dz * dz);
        }
        {
            const float dx = bb.x - px, dy = bb.y - py, dz = bb.z - pz;
            r2 = wn_max(r2, dx * dx + dy * dy + dz * dz);
        }
        {
            const float dx = c.x - px, dy = c.y - py, dz = c.z - pz;
            r2 = wn_max(r2, dx * dx + dy * dy + dz * dz);
        }
        // one ulp of slack upward so rounding can never make the radius smaller than a vertex distance
        r2 = r2 * 1.0000004f;
        WN_ATOMIC_MAX_U32(&b.r2v[node], (unsigned)wn_float_as_int(r2));
        node = b.parent[node];
        if (++guard > (1 << 20)) break;
    }
}

// ---- K5 : depth-first index by walking up, then pack -------------------------------------------------------------
WN_HD void wn_pack_node(const WnBuild& b, int node)
{
    int idx = 0, tf = 0, depth = 0;
    bool hidden = false;
    int cur = node;
    while (true) {
        const int p = b.parent[cur];
        if (p < 0) break;
        if (b.collapsed[p]) hidden = true;
        idx += (b.skip && b.skip[p]) ? 0 : 1;
        const int sl = b.slot[cur];
        for (int s = 0; s < sl; ++s) {
            const int c = b.child[(int64_t)p * b.W + s];
            if (c >= 0) {
                idx += b.size[c];
                tf += b.ntri[c];
            }
        }
        cur = p;
        if (++depth > (1 << 20)) {
            *b.err = WN_ERR_TOPOLOGY_DEPTH;
            return;
        }
    }
    if (cur != 0) { // not connected to the root
        *b.err = WN_ERR_TOPOLOGY_BAD_CHILD;
        return;
    }
    WN_ATOMIC_MAX_INT(b.max_depth, depth);
    const bool is_tri_leaf = node >= b.nI;
    if (is_tri_leaf) {
        const int t = (int)b.prim[node - b.nI];
        const WnV3 a = wn_vertex(b.v_xyz, b.tri[3 * (int64_t)t]);
        const WnV3 bb = wn_vertex(b.v_xyz, b.tri[3 * (int64_t)t + 1]);
        const WnV3 c = wn_vertex(b.v_xyz, b.tri[3 * (int64_t)t + 2]);
        b.tris[3 * (int64_t)tf] = make_float4(a.x, a.y, a.z, 0.0f);
        b.tris[3 * (int64_t)tf + 1] = make_float4(bb.x, bb.y, bb.z, 0.0f);
        b.tris[3 * (int64_t)tf + 2] = make_float4(c.x, c.y, c.z, 0.0f);
        b.tri_order[tf] = (unsigned)t;
    }
    if (hidden) return;
    if (!is_tri_leaf && b.skip && b.skip[node]) return; // no record: its children are its parent's children
    const bool leaf_entry = is_tri_leaf || b.collapsed[node];
    WnLocal d;
    wn_load_local(b.local + (int64_t)node * 9, d, false);
    float r2;
    const float inf = wn_int_as_float(0x7f800000);
    if (node == 0) {
        r2 = inf; // the reference never approximates the root (A.5)
    } else if (is_tri_leaf && !b.approx_single) {
        r2 = inf; // a lone triangle is cheaper exact than expanded
    } else if (b.radius_mode == 1) {
        r2 = wn_min(wn_int_as_float((int)b.r2v[node]), wn_box_corner_r2(d));
    } else {
        r2 = wn_box_corner_r2(d);
    }
    float4 rec[6];
    wn_pack_record(d, r2, leaf_entry, b.order, rec);
    const int link = leaf_entry ? wn_leaf_link(tf, b.ntri[node]) : idx + b.size[node];
    rec[1].w = wn_int_as_float(link); // raw bits, never used as a number
    b.hot[2 * (int64_t)idx] = rec[0];
    b.hot[2 * (int64_t)idx + 1] = rec[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) b.cold[4 * (int64_t)idx + k] = rec[2 + k];
    int kid[WN_MAX_WIDTH] = {-1, -1, -1, -1};
    if (!leaf_entry) {
        int cidx = idx + 1, n = 0;
        for (int s = 0; s < b.W; ++s) {
            const int c = b.child[(int64_t)node * b.W + s];
            if (c < 0) continue;
            if (c < b.nI && b.skip && b.skip[c] && !b.collapsed[c]) {
                // skipped child: its children (never skipped themselves) take its place
                for (int s2 = 0; s2 < b.W; ++s2) {
                    const int g = b.child[(int64_t)c * b.W + s2];
                    if (g < 0) continue;
                    if (n < WN_MAX_WIDTH) kid[n] = cidx;
                    ++n;
                    cidx += b.size[g];
                }
            } else {
                if (n < WN_MAX_WIDTH) kid[n] = cidx;
                ++n;
                cidx += b.size[c];
            }
        }
        if (n > WN_MAX_WIDTH) *b.err = WN_ERR_TOPOLOGY_BAD_CHILD;
    }
    b.kids[idx] = make_int4(kid[0], kid[1], kid[2], kid[3]);
}
