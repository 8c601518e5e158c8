// wn_refbuild_core.cuh — K3R: the reference builder's hierarchy (UT_BVH<4>::init<BOX_AREA>) built level by
// level on the GPU, so that wn_create itself — not an imported host tree — yields the tree the reference algorithm walks.
//
// Replaces the topology half of UT_SolidAngle::init (adobe/lagrange modules/winding/src/FastWindingNumber.cpp:57). The rule
// set is the one oracle/wn_oracle.cpp restates (BvhBuilder::split / multi_split / init_node): a range of more than four
// triangles becomes a 4-ary node by three successive binary splits (first the range, then twice the sub range with the
// largest half-area x count); a binary split is exhaustive for <= 6 items, a sorted sweep for <= 32, and otherwise the
// cheapest boundary of 16 equal-width spans along the longest axis of the range's box with a 1/16 balance guard and an
// order-statistic fallback. Every float operation is unfused and in the restatement's order, every partition is stable, so
// the topology (child table, DFS numbering, order of the triangle children) is the restatement's bit for bit
// (tests/test_emulation.py on the CPU tier, tests/test_gpu_reference_tree.py on the GPU).
//
// Data-parallel formulation. Top-down recursion becomes a loop over LEVELS of the 4-ary tree; inside a level every open node
// (range > 4 triangles) runs its three binary splits in three ROUNDS, all nodes in lock step:
//   round begin  (one thread per node)  pick the sub range to split; ranges <= 32 triangles are split right there, in place
//   bin          (one thread per item)  larger ranges: span boxes + counts with atomics (block-privatised in shared memory)
//   decide       (one thread per node)  cheapest balanced span boundary, or the order-statistic fallback
//   partition    (one thread per item)  stable: flag, global exclusive scan, scatter (other ranges are copied through)
//   fallback     (rare)                 one stable radix sort keyed (range start, centre) orders the affected ranges in place
//   emit         (one thread per node)  child rows, closed (<= 4 triangles) children, the next level's open nodes
// Node ids are assigned in creation order and renumbered at the end: the reference numbers nodes in depth-first pre-order,
// which is the order of (range start, depth) — the nodes that share a start form a chain of first children, so
//   id = #nodes starting left of it + (depth - depth of the chain's top)  — one scan over positions, no traversal.
//
// Written as host/device functors + a backend-templated driver: the sm_100a build runs them as kernels
// (wn_refbuild.cuh), tests/emul runs the same driver with a sequential backend on the CPU tier (test infrastructure).
#pragma once

#include "wn_build_core.cuh"

#define WN_REF_NSPANS 16
#define WN_REF_SMALL 6
#define WN_REF_MID 32
#define WN_REF_MINFRAC 16
#define WN_REF_ROW (WN_REF_NSPANS * 7) /* ints per span-table row: 16 boxes (6 ordered ints each) + 16 counts */
#define WN_REF_FLT_MAX 3.402823466e38f

#if defined(__CUDA_ARCH__)
#define WN_ATOMIC_MIN_INT(p, v) atomicMin((p), (v))
#define WN_ATOMIC_ADD_U32(p, v) atomicAdd((p), (v))
#else
static inline int wn_host_atomic_min(int* p, int v)
{
    const int o = *p;
    if (v < o) *p = v;
    return o;
}
static inline unsigned wn_host_atomic_add_u(unsigned* p, unsigned v)
{
    const unsigned o = *p;
    *p = o + v;
    return o;
}
#define WN_ATOMIC_MIN_INT(p, v) wn_host_atomic_min((p), (v))
#define WN_ATOMIC_ADD_U32(p, v) wn_host_atomic_add_u((p), (v))
#endif

// One 4-ary node under construction.
struct WnRefNode
{
    int sub[5];      // boundaries of its sub ranges (positions in the item order): sub[0] = start, sub[nsub] = end
    int nsub;        // 1 when the level starts, 4 when it ends
    int tmp_id;      // creation index = row in the temporary child table
    int depth;       // depth in the 4-ary tree (root 0)
    int chain_top;   // depth of the shallowest node that starts at the same position (see the header comment)
    int choice;      // sub range split in the current round
    int pad[2];
    float box[4][6]; // boxes of the sub ranges: lo xyz, hi xyz (union of the triangle boxes; the root: of all triangles)
};

// The binary split a node performs in the current round, as far as the per-item passes need it (32 bytes).
struct WnRefTask
{
    int ts, tn;      // the range being split
    int mode;        // 0: nothing for the items to do; 1: binned split; 2: order-statistic fallback
    int axis;
    float axis_min_x2, scale; // span of an item = clamp(int((lo + hi - axis_min_x2) * scale), 0, 15)
    int split_index; // binned: items with span <= split_index go left
    int nleft;       // left count (fallback: the order statistic)
};

struct WnRefState
{
    int N;
    const float4* tbox; // [2N] triangle boxes: (lo, -) (hi, -)
    unsigned* idx;      // [N] current item order (triangle ids)
    unsigned* idx_alt;  // [N]
    int* owner;         // [N] open node of the current level that holds the item, -1: none (its range is closed)
    uint32_t* flag;     // [N] partition flags / their scan
    WnRefNode* nodes;   // open nodes of the current level
    WnRefNode* next;    // open nodes of the next level
    WnRefTask* tasks;   // [count]
    int* rows;          // span tables, row = range start / 32
    uint32_t* copen;    // [count] children with more than 4 triangles (scanned)
    uint32_t* cnode;    // [count] children with more than 1 triangle (scanned)
    int* child_tmp;     // [4 * cap] child table in creation order
    int* info_start;    // [cap] per created node: range start, depth, chain top
    int* info_depth;
    int* info_chain;
    uint32_t* cnt_start; // [N] nodes starting at each position (scanned at the end)
    int* final_of;      // [cap] creation id -> reference id
    int* child_out;     // [4 * cap] final child table (wn_create_from_topology encoding)
    uint64_t* keys;     // [N] fallback sort keys
    int* res;           // [8] 0: a binned split exists, 1: a fallback exists, 2: open nodes of the next level, 3: nodes created so far,
                        //     4: largest range of the next level
    int count;          // open nodes of the current level
    int base_tmp;       // nodes created before this level's emit
    int r0, r1;         // rounds [r0, r1) handled by one WnRefRoundBegin pass (r1 - r0 > 1 only when no range exceeds 32)
};

WN_HD int wn_ref_ordered(float f)
{
    const int i = wn_float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
WN_HD float wn_ref_unordered(int i)
{
    if (i == 0x7fffffff) return WN_REF_FLT_MAX;         // untouched minimum
    if (i == (int)0x80000000) return -WN_REF_FLT_MAX;   // untouched maximum
    return wn_int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

WN_HD void wn_ref_tri_box(const float4* tbox, unsigned t, float* b)
{
    const float4 lo = tbox[2 * (size_t)t], hi = tbox[2 * (size_t)t + 1];
    b[0] = lo.x, b[1] = lo.y, b[2] = lo.z, b[3] = hi.x, b[4] = hi.y, b[5] = hi.z;
}
// std::min / std::max as the restatement uses them (the first argument survives ties and NaN second arguments)
WN_HD float wn_ref_min(float a, float b)
{
    return b < a ? b : a;
}
WN_HD float wn_ref_max(float a, float b)
{
    return a < b ? b : a;
}
WN_HD void wn_ref_combine(float* acc, const float* b)
{
    for (int a = 0; a < 3; ++a) {
        acc[a] = wn_ref_min(acc[a], b[a]);
        acc[3 + a] = wn_ref_max(acc[3 + a], b[3 + a]);
    }
}
WN_HD void wn_ref_copy_box(float* dst, const float* src)
{
    for (int a = 0; a < 6; ++a) dst[a] = src[a];
}
WN_HD void wn_ref_empty_box(float* b)
{
    b[0] = b[1] = b[2] = WN_REF_FLT_MAX;
    b[3] = b[4] = b[5] = -WN_REF_FLT_MAX;
}
// Box::half_surface_area of the restatement: d0 d1 + d1 d2 + d2 d0, unfused, left to right
WN_HD float wn_ref_half_area(const float* b)
{
    const float d0 = WN_SUB(b[3], b[0]), d1 = WN_SUB(b[4], b[1]), d2 = WN_SUB(b[5], b[2]);
    return WN_ADD(WN_ADD(WN_MUL(d0, d1), WN_MUL(d1, d2)), WN_MUL(d2, d0));
}
// longest axis of a box (first one on ties) and its length
WN_HD int wn_ref_axis(const float* b, float& len)
{
    int axis = 0;
    len = WN_SUB(b[3], b[0]);
    for (int a = 1; a < 3; ++a) {
        const float l = WN_SUB(b[3 + a], b[a]);
        if (l > len) {
            axis = a;
            len = l;
        }
    }
    return axis;
}
WN_HD int wn_ref_span(const WnRefTask& t, const float* box)
{
    const float sum = WN_ADD(box[t.axis], box[3 + t.axis]);
    const int s = (int)WN_MUL(WN_SUB(sum, t.axis_min_x2), t.scale);
    return s < 0 ? 0 : (s > WN_REF_NSPANS - 1 ? WN_REF_NSPANS - 1 : s);
}

// Binary split of a range of n <= 32 items (or of a range whose box has no extent), in place. Returns the left count and the
// boxes of the two sides; -1 if the range needs the binned path. Mirrors BvhBuilder::split (oracle/wn_oracle.cpp) case by case.
WN_HD int wn_ref_split_small(const float4* tbox, unsigned* idx, int n, const float* abox, float* lbox, float* rbox)
{
    if (n == 2) {
        wn_ref_tri_box(tbox, idx[0], lbox);
        wn_ref_tri_box(tbox, idx[1], rbox);
        return 1;
    }
    if (n <= WN_REF_SMALL) {
        // exhaustive search over the two-way partitions with item 0 on side 0; first minimum wins
        float local[WN_REF_SMALL][6];
        unsigned ids[WN_REF_SMALL];
        for (int i = 0; i < n; ++i) {
            ids[i] = idx[i];
            wn_ref_tri_box(tbox, ids[i], local[i]);
        }
        const int limit = 1 << (n - 1);
        int best_bits = -1;
        float best_h = 0.0f;
        for (int bits = 1; bits < limit; ++bits) {
            float sb[2][6];
            wn_ref_copy_box(sb[0], local[0]);
            wn_ref_empty_box(sb[1]);
            int cnt[2] = {1, 0};
            for (int b = 0; b < n - 1; ++b) {
                const int dest = (bits >> b) & 1;
                wn_ref_combine(sb[dest], local[b + 1]);
                ++cnt[dest];
            }
            const float h = WN_ADD(WN_MUL(wn_ref_half_area(sb[0]), (float)cnt[0]), WN_MUL(wn_ref_half_area(sb[1]), (float)cnt[1]));
            if (best_bits == -1 || h < best_h) {
                best_bits = bits;
                best_h = h;
                wn_ref_copy_box(lbox, sb[0]);
                wn_ref_copy_box(rbox, sb[1]);
            }
        }
        int k = 0;
        idx[k++] = ids[0];
        for (int b = 0; b < n - 1; ++b)
            if (!((best_bits >> b) & 1)) idx[k++] = ids[b + 1];
        const int nleft = k;
        for (int b = 0; b < n - 1; ++b)
            if ((best_bits >> b) & 1) idx[k++] = ids[b + 1];
        return nleft;
    }
    float axis_len;
    const int axis = wn_ref_axis(abox, axis_len);
    if (!(axis_len > 0.0f)) {
        // all boxes are one point (or NaN): arbitrary middle split, both sides keep the parent's box
        wn_ref_copy_box(lbox, abox);
        wn_ref_copy_box(rbox, abox);
        return n / 2;
    }
    if (n > WN_REF_MID) return -1;
    // stable insertion sort by the box centre along the axis, then every split position, scanned from the right (strict <)
    unsigned ids[WN_REF_MID];
    float key[WN_REF_MID];
    for (int i = 0; i < n; ++i) {
        const unsigned t = idx[i];
        float b[6];
        wn_ref_tri_box(tbox, t, b);
        const float kx = WN_ADD(b[axis], b[3 + axis]);
        int j = i;
        while (j > 0 && key[j - 1] > kx) {
            key[j] = key[j - 1];
            ids[j] = ids[j - 1];
            --j;
        }
        key[j] = kx;
        ids[j] = t;
    }
    for (int i = 0; i < n; ++i) idx[i] = ids[i];
    float left[WN_REF_MID - 1][6];
    wn_ref_tri_box(tbox, ids[0], left[0]);
    for (int i = 1; i < n - 1; ++i) {
        float b[6];
        wn_ref_tri_box(tbox, ids[i], b);
        wn_ref_copy_box(left[i], left[i - 1]);
        wn_ref_combine(left[i], b);
    }
    float right[6];
    wn_ref_tri_box(tbox, ids[n - 1], right);
    int best_left = n - 1;
    float best_h = WN_ADD(WN_MUL(wn_ref_half_area(left[n - 2]), (float)(n - 1)), WN_MUL(wn_ref_half_area(right), 1.0f));
    wn_ref_copy_box(lbox, left[n - 2]);
    wn_ref_copy_box(rbox, right);
    for (int lc = n - 2; lc > 0; --lc) {
        float b[6];
        wn_ref_tri_box(tbox, ids[lc], b);
        wn_ref_combine(right, b);
        const float h = WN_ADD(WN_MUL(wn_ref_half_area(left[lc - 1]), (float)lc), WN_MUL(wn_ref_half_area(right), (float)(n - lc)));
        if (h < best_h) {
            best_h = h;
            best_left = lc;
            wn_ref_copy_box(lbox, left[lc - 1]);
            wn_ref_copy_box(rbox, right);
        }
    }
    return best_left;
}

// multi_split's bookkeeping: sub range `c` of the node was split after its first nleft items
WN_HD void wn_ref_insert_split(WnRefNode& nd, int c, int nleft, const float* lbox, const float* rbox)
{
    for (int i = nd.nsub; i > c; --i) nd.sub[i + 1] = nd.sub[i];
    for (int i = nd.nsub - 1; i > c; --i) wn_ref_copy_box(nd.box[i + 1], nd.box[i]);
    nd.sub[c + 1] = nd.sub[c] + nleft;
    wn_ref_copy_box(nd.box[c], lbox);
    wn_ref_copy_box(nd.box[c + 1], rbox);
    ++nd.nsub;
}

// ---- per-node pass at the start of a round ---------------------------------------------------------------------------
struct WnRefRoundBegin
{
    WnRefState s;
    WN_HD void operator()(int64_t i) const
    {
        WnRefNode nd = s.nodes[i];
        WnRefTask task;
        task.ts = 0, task.tn = 0, task.mode = 0, task.axis = 0, task.axis_min_x2 = 0.0f, task.scale = 0.0f, task.split_index = -1, task.nleft = 0;
        for (int r = s.r0; r < s.r1; ++r) {
            // the sub range with the largest half-area x count among those with more than one item (first maximum)
            int choice = -1;
            float max_h = 0.0f;
            for (int k = 0; k < nd.nsub; ++k) {
                const int cnt = nd.sub[k + 1] - nd.sub[k];
                if (cnt > 1) {
                    const float h = WN_MUL(wn_ref_half_area(nd.box[k]), (float)cnt);
                    if (choice == -1 || h > max_h) {
                        choice = k;
                        max_h = h;
                    }
                }
            }
            if (r == 0) choice = 0; // the node's own range (multi_split's first split does not look at the measure)
            if (choice < 0) break;  // cannot happen for a range of more than four items
            nd.choice = choice;
            const int ts = nd.sub[choice], tn = nd.sub[choice + 1] - ts;
            float lbox[6], rbox[6];
            const int nleft = wn_ref_split_small(s.tbox, s.idx + ts, tn, nd.box[choice], lbox, rbox);
            if (nleft >= 0) {
                wn_ref_insert_split(nd, choice, nleft, lbox, rbox);
                continue;
            }
            // binned split: the items do the measuring (WnRefBin), WnRefDecide picks the boundary
            float axis_len;
            task.axis = wn_ref_axis(nd.box[choice], axis_len);
            task.ts = ts;
            task.tn = tn;
            task.mode = 1;
            task.axis_min_x2 = WN_MUL(nd.box[choice][task.axis], 2.0f);
            task.scale = WN_DIV((float)WN_REF_NSPANS, WN_MUL(axis_len, 2.0f));
            int* row = s.rows + (size_t)(ts / WN_REF_MID) * WN_REF_ROW;
            for (int k = 0; k < WN_REF_NSPANS * 6; ++k) row[k] = (k % 6) < 3 ? 0x7fffffff : (int)0x80000000;
            for (int k = 0; k < WN_REF_NSPANS; ++k) row[WN_REF_NSPANS * 6 + k] = 0;
            s.res[0] = 1;
        }
        s.nodes[i] = nd;
        s.tasks[i] = task;
    }
};

// ---- per-item pass: span boxes and counts of the binned splits ---------------------------------------------------------
struct WnRefBin
{
    WnRefState s;
    WN_HD void operator()(int64_t p) const
    {
        const int i = s.owner[p];
        if (i < 0) return;
        const WnRefTask t = s.tasks[i];
        if (t.mode != 1 || p < t.ts || p >= t.ts + t.tn) return;
        float b[6];
        wn_ref_tri_box(s.tbox, s.idx[p], b);
        const int sp = wn_ref_span(t, b);
        int* row = s.rows + (size_t)(t.ts / WN_REF_MID) * WN_REF_ROW;
        for (int a = 0; a < 3; ++a) {
            WN_ATOMIC_MIN_INT(&row[sp * 6 + a], wn_ref_ordered(b[a]));
            WN_ATOMIC_MAX_INT(&row[sp * 6 + 3 + a], wn_ref_ordered(b[3 + a]));
        }
        WN_ATOMIC_ADD_INT(&row[WN_REF_NSPANS * 6 + sp], 1);
    }
};

// ---- per-node pass: the cheapest balanced span boundary, or the order-statistic fallback -------------------------------
struct WnRefDecide
{
    WnRefState s;
    WN_HD void operator()(int64_t i) const
    {
        WnRefTask t = s.tasks[i];
        if (t.mode != 1) return;
        int* row = s.rows + (size_t)(t.ts / WN_REF_MID) * WN_REF_ROW;
        float span[WN_REF_NSPANS][6];
        int cnt[WN_REF_NSPANS];
        for (int k = 0; k < WN_REF_NSPANS; ++k) {
            for (int a = 0; a < 6; ++a) span[k][a] = wn_ref_unordered(row[k * 6 + a]);
            cnt[k] = row[WN_REF_NSPANS * 6 + k];
        }
        float rbx[WN_REF_NSPANS - 1][6]; // boxes of spans k+1 .. 15
        {
            float acc[6];
            wn_ref_copy_box(acc, span[WN_REF_NSPANS - 1]);
            wn_ref_copy_box(rbx[WN_REF_NSPANS - 2], acc);
            for (int k = WN_REF_NSPANS - 3; k >= 0; --k) {
                wn_ref_combine(acc, span[k + 1]);
                wn_ref_copy_box(rbx[k], acc);
            }
        }
        const int n = t.tn;
        const int min_count = n / WN_REF_MINFRAC;
        const int max_count = (int)(((long long)(WN_REF_MINFRAC - 1) * (long long)n) / WN_REF_MINFRAC);
        float smallest = wn_int_as_float(0x7f800000);
        int split_index = -1, best_lc = 0, lc = 0, lc0 = 0, lc_last = 0;
        float acc[6], lbox[6], rbox[6];
        wn_ref_empty_box(acc);
        for (int k = 0; k < WN_REF_NSPANS - 1; ++k) {
            wn_ref_combine(acc, span[k]);
            lc += cnt[k];
            if (k == 0) lc0 = lc;
            lc_last = lc;
            if (lc < min_count || lc > max_count) continue;
            const float h = WN_ADD(WN_MUL((float)lc, wn_ref_half_area(acc)), WN_MUL((float)(n - lc), wn_ref_half_area(rbx[k])));
            if (h < smallest) {
                smallest = h;
                split_index = k;
                best_lc = lc;
                wn_ref_copy_box(lbox, acc);
                wn_ref_copy_box(rbox, rbx[k]);
            }
        }
        if (split_index >= 0) {
            t.split_index = split_index;
            t.nleft = best_lc;
            s.tasks[i] = t;
            WnRefNode nd = s.nodes[i];
            wn_ref_insert_split(nd, nd.choice, best_lc, lbox, rbox);
            s.nodes[i] = nd;
            return;
        }
        // nothing balanced: an order statistic of the centres instead. The range is ordered by one stable sort (WnRefFbKeys);
        // its two boxes are measured afterwards (WnRefFbBoxes) into spans 0 and 1 of the row.
        t.mode = 2;
        t.nleft = lc0 > max_count ? max_count : (lc_last < min_count ? min_count : n / 2);
        s.tasks[i] = t;
        for (int k = 0; k < 12; ++k) row[k] = (k % 6) < 3 ? 0x7fffffff : (int)0x80000000;
        s.res[1] = 1;
    }
};

// ---- stable partition of the binned ranges: flag, (scan), scatter -------------------------------------------------------
struct WnRefFlag
{
    WnRefState s;
    WN_HD void operator()(int64_t p) const
    {
        uint32_t f = 0;
        const int i = s.owner[p];
        if (i >= 0) {
            const WnRefTask t = s.tasks[i];
            if (t.mode == 1 && p >= t.ts && p < t.ts + t.tn) {
                float b[6];
                wn_ref_tri_box(s.tbox, s.idx[p], b);
                f = wn_ref_span(t, b) <= t.split_index ? 1u : 0u;
            }
        }
        s.flag[p] = f;
    }
};
struct WnRefScatter
{
    WnRefState s; // flag holds the exclusive scan of the flags
    WN_HD void operator()(int64_t p) const
    {
        const unsigned tri = s.idx[p];
        int64_t dst = p;
        const int i = s.owner[p];
        if (i >= 0) {
            const WnRefTask t = s.tasks[i];
            if (t.mode == 1 && p >= t.ts && p < t.ts + t.tn) {
                float b[6];
                wn_ref_tri_box(s.tbox, tri, b);
                const bool left = wn_ref_span(t, b) <= t.split_index;
                const int rank_left = (int)(s.flag[p] - s.flag[t.ts]);
                dst = left ? t.ts + rank_left : t.ts + t.nleft + ((int)(p - t.ts) - rank_left);
            }
        }
        s.idx_alt[dst] = tri;
    }
};

// ---- fallback: order the affected ranges by centre (stable), measure the two sides ---------------------------------------
struct WnRefFbKeys
{
    WnRefState s;
    WN_HD void operator()(int64_t p) const
    {
        uint64_t key = (uint64_t)p << 32; // everything else keeps its position
        const int i = s.owner[p];
        if (i >= 0) {
            const WnRefTask t = s.tasks[i];
            if (t.mode == 2 && p >= t.ts && p < t.ts + t.tn) {
                float b[6];
                wn_ref_tri_box(s.tbox, s.idx[p], b);
                float c = WN_ADD(b[t.axis], b[3 + t.axis]);
                if (c == 0.0f) c = 0.0f; // -0 and +0 compare equal in the reference's comparator
                const uint32_t u = (uint32_t)wn_float_as_int(c);
                key = ((uint64_t)t.ts << 32) | (uint64_t)(u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u));
            }
        }
        s.keys[p] = key;
    }
};
struct WnRefFbBoxes
{
    WnRefState s;
    WN_HD void operator()(int64_t p) const
    {
        const int i = s.owner[p];
        if (i < 0) return;
        const WnRefTask t = s.tasks[i];
        if (t.mode != 2 || p < t.ts || p >= t.ts + t.tn) return;
        float b[6];
        wn_ref_tri_box(s.tbox, s.idx[p], b);
        const int sp = (p - t.ts) < t.nleft ? 0 : 1;
        int* row = s.rows + (size_t)(t.ts / WN_REF_MID) * WN_REF_ROW;
        for (int a = 0; a < 3; ++a) {
            WN_ATOMIC_MIN_INT(&row[sp * 6 + a], wn_ref_ordered(b[a]));
            WN_ATOMIC_MAX_INT(&row[sp * 6 + 3 + a], wn_ref_ordered(b[3 + a]));
        }
    }
};
struct WnRefFbFinish
{
    WnRefState s;
    WN_HD void operator()(int64_t i) const
    {
        const WnRefTask t = s.tasks[i];
        if (t.mode != 2) return;
        const int* row = s.rows + (size_t)(t.ts / WN_REF_MID) * WN_REF_ROW;
        float lbox[6], rbox[6];
        for (int a = 0; a < 6; ++a) {
            lbox[a] = wn_ref_unordered(row[a]);
            rbox[a] = wn_ref_unordered(row[6 + a]);
        }
        WnRefNode nd = s.nodes[i];
        wn_ref_insert_split(nd, nd.choice, t.nleft, lbox, rbox);
        s.nodes[i] = nd;
    }
};

// ---- end of a level: child rows, closed children, the next level's open nodes ----------------------------------------------
struct WnRefCount
{
    WnRefState s;
    WN_HD void operator()(int64_t i) const
    {
        const WnRefNode& nd = s.nodes[i];
        uint32_t open = 0, node = 0;
        for (int k = 0; k < nd.nsub; ++k) {
            const int m = nd.sub[k + 1] - nd.sub[k];
            open += m > 4 ? 1u : 0u;
            node += m > 1 ? 1u : 0u;
        }
        s.copen[i] = open;
        s.cnode[i] = node;
    }
};
struct WnRefEmit
{
    WnRefState s; // copen / cnode hold their exclusive scans
    WN_HD void operator()(int64_t i) const
    {
        const WnRefNode nd = s.nodes[i];
        int* row = s.child_tmp + 4 * (size_t)nd.tmp_id;
        uint32_t ko = s.copen[i], kn = s.cnode[i];
        int max_n = 0;
        for (int k = 0; k < 4; ++k) {
            if (k >= nd.nsub) {
                row[k] = WN_CHILD_EMPTY;
                continue;
            }
            const int st = nd.sub[k], m = nd.sub[k + 1] - st;
            if (m == 1) {
                row[k] = wn_enc_tri((int)s.idx[st]);
                continue;
            }
            const int cid = s.base_tmp + (int)kn++;
            row[k] = cid;
            s.info_start[cid] = st;
            s.info_depth[cid] = nd.depth + 1;
            s.info_chain[cid] = k == 0 ? nd.chain_top : nd.depth + 1;
            WN_ATOMIC_ADD_U32(&s.cnt_start[st], 1u);
            if (m <= 4) {
                int* crow = s.child_tmp + 4 * (size_t)cid;
                for (int j = 0; j < 4; ++j) crow[j] = j < m ? wn_enc_tri((int)s.idx[st + j]) : WN_CHILD_EMPTY;
            } else {
                WnRefNode c;
                c.sub[0] = st, c.sub[1] = st + m, c.sub[2] = c.sub[3] = c.sub[4] = 0;
                c.nsub = 1;
                c.tmp_id = cid;
                c.depth = nd.depth + 1;
                c.chain_top = k == 0 ? nd.chain_top : nd.depth + 1;
                c.choice = 0;
                c.pad[0] = c.pad[1] = 0;
                wn_ref_copy_box(c.box[0], nd.box[k]);
                for (int j = 1; j < 4; ++j) wn_ref_empty_box(c.box[j]);
                s.next[ko++] = c;
                max_n = m > max_n ? m : max_n;
            }
        }
        if (max_n > 0) WN_ATOMIC_MAX_INT(&s.res[4], max_n);
        if (i == s.count - 1) {
            s.res[2] = (int)ko;
            s.res[3] = s.base_tmp + (int)kn;
        }
    }
};
struct WnRefOwner
{
    WnRefState s; // copen holds its exclusive scan
    WN_HD void operator()(int64_t p) const
    {
        const int i = s.owner[p];
        if (i < 0) return;
        const WnRefNode& nd = s.nodes[i];
        int o = (int)s.copen[i], res = -1;
        for (int k = 0; k < nd.nsub; ++k) {
            const int m = nd.sub[k + 1] - nd.sub[k];
            if (p < nd.sub[k + 1]) {
                res = m > 4 ? o : -1;
                break;
            }
            o += m > 4 ? 1 : 0;
        }
        s.owner[p] = res;
    }
};

// ---- the reference's numbering (depth-first pre-order) and the final table -----------------------------------------------------
struct WnRefNumber
{
    WnRefState s; // cnt_start holds its exclusive scan
    WN_HD void operator()(int64_t c) const { s.final_of[c] = (int)s.cnt_start[s.info_start[c]] + s.info_depth[c] - s.info_chain[c]; }
};
struct WnRefRemap
{
    WnRefState s;
    WN_HD void operator()(int64_t c) const
    {
        const int f = s.final_of[c];
        for (int k = 0; k < 4; ++k) {
            const int v = s.child_tmp[4 * (size_t)c + k];
            s.child_out[4 * (size_t)f + k] = v >= 0 ? s.final_of[v] : v;
        }
    }
};

struct WnRefTriBoxes
{
    const float* v_xyz;
    const int* tri;
    float4* tbox;
    unsigned* idx;
    int* owner;
    WN_HD void operator()(int64_t t) const
    {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            const float x0 = v_xyz[3 * (size_t)tri[3 * t] + a], x1 = v_xyz[3 * (size_t)tri[3 * t + 1] + a], x2 = v_xyz[3 * (size_t)tri[3 * t + 2] + a];
            lo[a] = wn_ref_min(wn_ref_min(wn_ref_min(WN_REF_FLT_MAX, x0), x1), x2); // the restatement's init_empty + three combines
            hi[a] = wn_ref_max(wn_ref_max(wn_ref_max(-WN_REF_FLT_MAX, x0), x1), x2);
        }
        tbox[2 * t] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        tbox[2 * t + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
        idx[t] = (unsigned)t;
        owner[t] = 0;
    }
};

// the root: node 0 over all items, its box = union of all triangle boxes (rows[0..5], ordered ints, from Backend::root_bounds)
struct WnRefInitRoot
{
    WnRefState s;
    WN_HD void operator()(int64_t) const
    {
        WnRefNode root;
        root.sub[0] = 0, root.sub[1] = s.N, root.sub[2] = root.sub[3] = root.sub[4] = 0;
        root.nsub = 1;
        root.tmp_id = 0;
        root.depth = 0;
        root.chain_top = 0;
        root.choice = 0;
        root.pad[0] = root.pad[1] = 0;
        for (int a = 0; a < 6; ++a) root.box[0][a] = wn_ref_unordered(s.rows[a]);
        for (int j = 1; j < 4; ++j) wn_ref_empty_box(root.box[j]);
        s.nodes[0] = root;
        s.info_start[0] = 0;
        s.info_depth[0] = 0;
        s.info_chain[0] = 0;
        s.cnt_start[0] = 1u;
    }
};

// Per-N scratch the driver needs, in bytes (the caller allocates; see wn_ref_assign).
inline size_t wn_ref_align(size_t v)
{
    return (v + 255) / 256 * 256;
}
struct WnRefLayout
{
    size_t tbox, idx, idx_alt, owner, flag, nodes, next, tasks, rows, copen, cnode, child_tmp, info_start, info_depth, info_chain, cnt_start,
        final_of, keys, keys_alt, res, scan_scratch, total;
};
inline WnRefLayout wn_ref_layout(int64_t N, size_t scan_scratch_bytes)
{
    WnRefLayout L;
    size_t off = 0;
    const size_t open_cap = (size_t)N / 5 + 2, cap = (size_t)N + 2, rows = (size_t)N / WN_REF_MID + 2;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += wn_ref_align(bytes ? bytes : 1);
        return o;
    };
    L.tbox = take((size_t)N * 2 * sizeof(float4));
    L.idx = take((size_t)N * 4);
    L.idx_alt = take((size_t)N * 4);
    L.owner = take((size_t)N * 4);
    L.flag = take((size_t)N * 4);
    L.nodes = take(open_cap * sizeof(WnRefNode));
    L.next = take(open_cap * sizeof(WnRefNode));
    L.tasks = take(open_cap * sizeof(WnRefTask));
    L.rows = take(rows * WN_REF_ROW * 4);
    L.copen = take(open_cap * 4);
    L.cnode = take(open_cap * 4);
    L.child_tmp = take(cap * 16);
    L.info_start = take(cap * 4);
    L.info_depth = take(cap * 4);
    L.info_chain = take(cap * 4);
    L.cnt_start = take((size_t)N * 4);
    L.final_of = take(cap * 4);
    L.keys = take((size_t)N * 8);
    L.keys_alt = take((size_t)N * 8);
    L.res = take(64);
    L.scan_scratch = take(scan_scratch_bytes);
    L.total = off;
    return L;
}

// The driver. Backend: for_each(n, functor), bin(state), root_bounds(state), scan(uint32*, n), sort64(keys, vals, vals_alt, n,
// end_bit) -> 0/1 (which value buffer holds the result), read(host int*, device int*, count) [synchronises], zero(ptr, bytes),
// write(device, host, bytes).
// On return: child_out holds *num_nodes rows (wn_create_from_topology encoding).
template <class Backend>
inline bool wn_ref_build_topology(Backend& B, WnRefState s, int* num_nodes, int* num_levels)
{
    const int N = s.N;
    *num_levels = 0;
    if (N <= 4) {
        // a single node of triangle children (init_node's n <= BVH_N case)
        int row[4];
        for (int j = 0; j < 4; ++j) row[j] = j < N ? wn_enc_tri(j) : WN_CHILD_EMPTY;
        B.write(s.child_out, row, sizeof(row));
        *num_nodes = 1;
        return true;
    }
    B.zero(s.cnt_start, (size_t)N * 4);
    B.root_bounds(s);
    B.for_each(1, WnRefInitRoot{s});
    s.count = 1;
    s.base_tmp = 1;
    int max_n = N, level = 0;
    while (s.count > 0) {
        if (++level > 4096) return false;
        if (max_n <= WN_REF_MID) {
            // every split of this level is small: the three rounds in one pass, no item passes
            s.r0 = 0, s.r1 = 3;
            B.for_each(s.count, WnRefRoundBegin{s});
        } else {
            for (int r = 0; r < 3; ++r) {
                s.r0 = r, s.r1 = r + 1;
                B.zero(s.res, 8);
                B.for_each(s.count, WnRefRoundBegin{s});
                B.bin(s);
                B.for_each(s.count, WnRefDecide{s});
                int h[2] = {0, 0};
                B.read(h, s.res, 2);
                if (!h[0]) continue; // no binned split in this round
                if (h[1]) {
                    int bits = 0;
                    while (((int64_t)1 << bits) < N) ++bits;
                    B.for_each(N, WnRefFbKeys{s});
                    if (B.sort64(s.keys, s.idx, s.idx_alt, N, 32 + bits)) {
                        unsigned* t = s.idx;
                        s.idx = s.idx_alt;
                        s.idx_alt = t;
                    }
                    B.for_each(N, WnRefFbBoxes{s});
                    B.for_each(s.count, WnRefFbFinish{s});
                }
                B.for_each(N, WnRefFlag{s});
                B.scan(s.flag, N);
                B.for_each(N, WnRefScatter{s});
                unsigned* t = s.idx;
                s.idx = s.idx_alt;
                s.idx_alt = t;
            }
        }
        B.for_each(s.count, WnRefCount{s});
        B.scan(s.copen, s.count);
        B.scan(s.cnode, s.count);
        B.zero(s.res + 2, 12);
        B.for_each(s.count, WnRefEmit{s});
        B.for_each(N, WnRefOwner{s});
        int h[3] = {0, 0, 0};
        B.read(h, s.res + 2, 3);
        WnRefNode* t = s.nodes;
        s.nodes = s.next;
        s.next = t;
        s.count = h[0];
        s.base_tmp = h[1];
        max_n = h[2];
    }
    const int total = s.base_tmp;
    B.scan(s.cnt_start, N);
    B.for_each(total, WnRefNumber{s});
    B.for_each(total, WnRefRemap{s});
    *num_nodes = total;
    *num_levels = level;
    return true;
}
