// wn_oracle.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into or called by the product path.
//
// PARITY UNPINNED: the arithmetic of lagrange::winding::FastWindingNumber lives in a third-party library that is
// NOT in /root/reference (HDK_Sample::UT_SolidAngle<float,float>, fetched by CPM from
// jdumas/WindingNumber @ a48b8f555b490afe7aab9159c7daaf83fa2cdf8e, cmake/recipes/external/winding_number.cmake:21-26),
// and no reference test holds a golden value for it (modules/winding/tests/test_fast_winding_number.cpp:80-83 is a
// SUCCEED() no-op). This file is therefore a *restatement of the published algorithm* (Barill et al., "Fast Winding
// Numbers for Soups and Clouds", SIGGRAPH 2018, cited at modules/volume/include/lagrange/volume/mesh_to_volume.h:34)
// as the HDK sample implements it, anchored on the reference's own call sites:
//   - adapter semantics (V -> float, F -> int, init(), computeSolidAngle(q)):  modules/winding/src/FastWindingNumber.cpp:31-76
//   - is_inside predicate  Omega / (4.f * pi) > 0.5f  evaluated in double:      modules/winding/src/FastWindingNumber.cpp:66
//                                                                               modules/core/include/lagrange/internal/constants.h:16
//   - engine semantics (UT_BVH<4> BOX_AREA build, order-2 moments, beta = 2):   SURVEY.md Appendix A (A.1 - A.6)
// Details marked (+) are recollection of upstream that cannot be verified in this container; they affect tree
// topology / last-bit rounding only, never the mathematical definition.
//
// Two oracles:
//   wno_exact64_*  : tree-independent ground truth, double precision sum of Van Oosterom-Strackee solid angles.
//   wno_ref_*      : the reference-algorithm restatement, float32, 4-ary SAH BVH, one triangle per leaf, order-2
//                    Taylor far field, accuracy scale beta, with counters T/A/E that define algorithmic flops.
//
// Built by oracle/Makefile into oracle/liboracle_wn.so and loaded (ctypes) only from tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs.

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Small vector helpers (float, no FMA contraction: the file is compiled with -ffp-contract=off to mirror the
// reference's ISO C++17 / no -march / no fast-math build, CMakeLists.txt:270-272).
// ---------------------------------------------------------------------------------------------------------------
struct V3
{
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length2(V3 a) { return dot(a, a); }
inline V3 vmax(V3 a, V3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }

struct Box
{
    float lo[3], hi[3];
    void init_empty()
    {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::numeric_limits<float>::max();
            hi[a] = -std::numeric_limits<float>::max();
        }
    }
    void combine(const Box& b)
    {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    // UT::Box::half_surface_area() for 3 axes (+): d0*d1 + d1*d2 + d2*d0  -- the BOX_AREA heuristic.
    float half_area() const
    {
        const float d0 = hi[0] - lo[0], d1 = hi[1] - lo[1], d2 = hi[2] - lo[2];
        return d0 * d1 + d1 * d2 + d2 * d0;
    }
    float center_x2(int axis) const { return lo[axis] + hi[axis]; }
};

// Neutral child encoding shared with the C-ABI's wn_create_from_topology (include/wn_b200.h):
//   c >= 0  : internal node index;  c == -1 : empty slot;  c <= -2 : triangle index -(c+2).
constexpr int32_t kEmpty = -1;
inline int32_t enc_tri(int32_t t) { return -(t + 2); }
inline bool is_tri(int32_t c) { return c <= -2; }
inline int32_t dec_tri(int32_t c) { return -(c + 2); }

constexpr int BVH_N = 4;

// The 4 child lanes of a node are evaluated 4-wide (SSE), like upstream's v4uf code path; 0 selects the scalar lane loop
// (same operations in the same order: the two are bit-identical, tests/test_oracle.py checks it).
int g_simd_lanes = 1;

struct Node
{
    int32_t child[BVH_N];
};

// ---------------------------------------------------------------------------------------------------------------
// UT_BVH<4>::init<BOX_AREA> restated. Top-down, max one item per leaf slot.
// ---------------------------------------------------------------------------------------------------------------
struct BvhBuilder
{
    const Box* boxes;
    std::vector<Node>* nodes;

    static constexpr int NSPANS = 16;
    static constexpr int NSPLITS = NSPANS - 1;
    static constexpr int SMALL_LIMIT = 6;
    static constexpr int MID_LIMIT = 2 * NSPANS;
    static constexpr int MIN_FRACTION = 16;

    // Split [indices, indices+n) in two; returns the split position and the two sub boxes (+).
    int split(const Box& axes_minmax, int32_t* indices, int n, Box split_boxes[2]) const
    {
        if (n == 2) {
            split_boxes[0] = boxes[indices[0]];
            split_boxes[1] = boxes[indices[1]];
            return 1;
        }
        if (n <= SMALL_LIMIT) {
            // Exhaustive search over the 2^(n-1)-1 two-way partitions with item 0 fixed on side 0.
            Box local[SMALL_LIMIT];
            for (int i = 0; i < n; ++i) local[i] = boxes[indices[i]];
            const int limit = 1 << (n - 1);
            int best_bits = -1;
            float best_h = 0;
            for (int bits = 1; bits < limit; ++bits) {
                Box sb[2];
                sb[0] = local[0];
                sb[1].init_empty();
                int cnt[2] = {1, 0};
                for (int b = 0; b < n - 1; ++b) {
                    const int dest = (bits >> b) & 1;
                    sb[dest].combine(local[b + 1]);
                    ++cnt[dest];
                }
                const float h = sb[0].half_area() * cnt[0] + sb[1].half_area() * cnt[1];
                if (best_bits == -1 || h < best_h) {
                    best_bits = bits;
                    best_h = h;
                    split_boxes[0] = sb[0];
                    split_boxes[1] = sb[1];
                }
            }
            int32_t tmp[SMALL_LIMIT];
            int k = 0;
            tmp[k++] = indices[0];
            for (int b = 0; b < n - 1; ++b)
                if (!((best_bits >> b) & 1)) tmp[k++] = indices[b + 1];
            const int nleft = k;
            for (int b = 0; b < n - 1; ++b)
                if ((best_bits >> b) & 1) tmp[k++] = indices[b + 1];
            for (int i = 0; i < n; ++i) indices[i] = tmp[i];
            return nleft;
        }

        int axis = 0;
        float axis_len = axes_minmax.hi[0] - axes_minmax.lo[0];
        for (int a = 1; a < 3; ++a) {
            const float l = axes_minmax.hi[a] - axes_minmax.lo[a];
            if (l > axis_len) {
                axis = a;
                axis_len = l;
            }
        }
        if (!(axis_len > 0.0f)) {
            // All boxes are one point (or NaN): arbitrary middle split.
            split_boxes[0] = axes_minmax;
            split_boxes[1] = axes_minmax;
            return n / 2;
        }

        if (n <= MID_LIMIT) {
            // Stable sort by box centre along the axis, then try every split (+: scan from the right, strict <).
            std::stable_sort(indices, indices + n, [&](int32_t a, int32_t b) {
                return boxes[a].center_x2(axis) < boxes[b].center_x2(axis);
            });
            Box left[MID_LIMIT];
            left[0] = boxes[indices[0]];
            for (int i = 1; i < n - 1; ++i) {
                left[i] = left[i - 1];
                left[i].combine(boxes[indices[i]]);
            }
            Box right = boxes[indices[n - 1]];
            int best_left = n - 1;
            float best_h = left[n - 2].half_area() * float(n - 1) + right.half_area() * 1.0f;
            split_boxes[0] = left[n - 2];
            split_boxes[1] = right;
            for (int left_count = n - 2; left_count > 0; --left_count) {
                right.combine(boxes[indices[left_count]]);
                const float h = left[left_count - 1].half_area() * float(left_count) +
                                right.half_area() * float(n - left_count);
                if (h < best_h) {
                    best_h = h;
                    best_left = left_count;
                    split_boxes[0] = left[left_count - 1];
                    split_boxes[1] = right;
                }
            }
            return best_left;
        }

        // Binned SAH: 16 spans along the longest axis of the node's box, box centres decide the span.
        Box span_boxes[NSPANS];
        int span_counts[NSPANS];
        for (int i = 0; i < NSPANS; ++i) {
            span_boxes[i].init_empty();
            span_counts[i] = 0;
        }
        const float axis_min_x2 = axes_minmax.lo[axis] * 2;
        const float axis_index_scale = float(NSPANS) / (axis_len * 2);
        auto span_of = [&](int32_t item) {
            const float sum = boxes[item].center_x2(axis);
            const int s = int((sum - axis_min_x2) * axis_index_scale);
            return std::min(std::max(s, 0), NSPANS - 1);
        };
        for (int i = 0; i < n; ++i) {
            const int s = span_of(indices[i]);
            ++span_counts[s];
            span_boxes[s].combine(boxes[indices[i]]);
        }
        Box left_boxes[NSPLITS], right_boxes[NSPLITS];
        Box acc = span_boxes[0];
        left_boxes[0] = acc;
        for (int i = 1; i < NSPLITS; ++i) {
            acc.combine(span_boxes[i]);
            left_boxes[i] = acc;
        }
        acc = span_boxes[NSPANS - 1];
        right_boxes[NSPLITS - 1] = acc;
        for (int i = NSPLITS - 2; i >= 0; --i) {
            acc.combine(span_boxes[i + 1]);
            right_boxes[i] = acc;
        }
        int left_counts[NSPLITS];
        int cacc = span_counts[0];
        left_counts[0] = cacc;
        for (int i = 1; i < NSPLITS; ++i) {
            cacc += span_counts[i];
            left_counts[i] = cacc;
        }
        // Balance guard: at least 1/16 of the items on each side.
        const int min_count = n / MIN_FRACTION;
        const int max_count = int((uint64_t(MIN_FRACTION - 1) * uint64_t(n)) / MIN_FRACTION);
        float smallest = std::numeric_limits<float>::infinity();
        int split_index = -1;
        for (int s = 0; s < NSPLITS; ++s) {
            const int lc = left_counts[s];
            if (lc < min_count || lc > max_count) continue;
            const int rc = n - lc;
            const float h = float(lc) * left_boxes[s].half_area() + float(rc) * right_boxes[s].half_area();
            if (h < smallest) {
                smallest = h;
                split_index = s;
            }
        }
        if (split_index == -1) {
            // Nothing balanced: select an order statistic of the centres instead.
            int nth;
            if (left_counts[0] > max_count)
                nth = max_count;
            else if (left_counts[NSPLITS - 1] < min_count)
                nth = min_count;
            else
                nth = n / 2;
            // (+) upstream calls std::nth_element here, whose permutation is implementation-defined (any arrangement with
            // the nth centre in place, smaller-or-equal ones before it and larger-or-equal ones after it is conforming).
            // This restatement fixes ONE conforming outcome, the stable-sorted order, so that the tree is a function of the
            // input alone and a data-parallel builder can reproduce it (lagrange_b200/csrc/wn_refbuild_core.cuh does).
            std::stable_sort(indices, indices + n, [&](int32_t a, int32_t b) {
                return boxes[a].center_x2(axis) < boxes[b].center_x2(axis);
            });
            Box lb = boxes[indices[0]];
            for (int i = 1; i < nth; ++i) lb.combine(boxes[indices[i]]);
            Box rb = boxes[indices[nth]];
            for (int i = nth + 1; i < n; ++i) rb.combine(boxes[indices[i]]);
            split_boxes[0] = lb;
            split_boxes[1] = rb;
            return nth;
        }
        // Partition by span (+: upstream partitions by the pivot coordinate and repairs round-off mismatches; the
        // span index is the quantity its counts and boxes were computed from, so partitioning on it is consistent).
        std::stable_partition(indices, indices + n, [&](int32_t item) { return span_of(item) <= split_index; });
        split_boxes[0] = left_boxes[split_index];
        split_boxes[1] = right_boxes[split_index];
        return left_counts[split_index];
    }

    // multiSplit for N = 4: split in two, then keep splitting the sub range with the largest area * count.
    void multi_split(const Box& axes_minmax, int32_t* indices, int n, int32_t* sub[BVH_N + 1], Box sub_boxes[BVH_N]) const
    {
        sub[0] = indices;
        sub[2] = indices + n;
        {
            Box sb[2];
            const int s = split(axes_minmax, indices, n, sb);
            sub[1] = indices + s;
            sub_boxes[0] = sb[0];
            sub_boxes[1] = sb[1];
        }
        int nsub = 2;
        while (nsub < BVH_N) {
            int choice = -1;
            float max_h = 0;
            for (int i = 0; i < nsub; ++i) {
                const int cnt = int(sub[i + 1] - sub[i]);
                if (cnt > 1) {
                    const float h = sub_boxes[i].half_area() * float(cnt);
                    if (choice == -1 || h > max_h) {
                        choice = i;
                        max_h = h;
                    }
                }
            }
            int32_t* start = sub[choice];
            const int cnt = int(sub[choice + 1] - start);
            for (int i = nsub; i > choice; --i) sub[i + 1] = sub[i];
            for (int i = nsub - 1; i > choice; --i) sub_boxes[i + 1] = sub_boxes[i];
            Box sb[2];
            const int s = split(sub_boxes[choice], start, cnt, sb);
            sub[choice + 1] = start + s;
            sub_boxes[choice] = sb[0];
            sub_boxes[choice + 1] = sb[1];
            ++nsub;
        }
    }

    void init_node(int nodei, const Box& axes_minmax, int32_t* indices, int n) const
    {
        if (n <= BVH_N) {
            Node nd;
            for (int i = 0; i < n; ++i) nd.child[i] = enc_tri(indices[i]);
            for (int i = n; i < BVH_N; ++i) nd.child[i] = kEmpty;
            (*nodes)[nodei] = nd;
            return;
        }
        int32_t* sub[BVH_N + 1];
        Box sub_boxes[BVH_N];
        multi_split(axes_minmax, indices, n, sub, sub_boxes);
        for (int i = 0; i < BVH_N; ++i) {
            const int cnt = int(sub[i + 1] - sub[i]);
            if (cnt == 1) {
                (*nodes)[nodei].child[i] = enc_tri(sub[i][0]);
            } else {
                const int child_node = int(nodes->size());
                nodes->push_back(Node{{kEmpty, kEmpty, kEmpty, kEmpty}});
                (*nodes)[nodei].child[i] = child_node;
                init_node(child_node, sub_boxes[i], sub[i], cnt);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Order-2 moments.
// ---------------------------------------------------------------------------------------------------------------
struct LocalData
{
    Box box;
    V3 avgP;
    V3 areaP;
    V3 N;
    float area;
    // full first-order tensor (needed to shift second order)
    V3 NijDiag; // Nxx, Nyy, Nzz
    float Nxy, Nyx, Nyz, Nzy, Nzx, Nxz;
    // second order, in the 10 combinations the query needs (closed under the shift given the 9 Nij)
    V3 NijkDiag; // Nxxx, Nyyy, Nzzz
    float SumPermuteNxyz; // 2 (Nxyz + Nyzx + Nzxy)
    float N2xxy_yxx, N2xxz_zxx, N2yyz_zyy, N2yyx_xyy, N2zzx_xzz, N2zzy_yzz;
};

// What the query reads for one child lane: 23 floats (A.4), in this fixed order (also the order of
// wno_ref_get_boxdata and of the C-ABI's debug moment dump).
enum
{
    F_PX = 0, F_PY, F_PZ, F_R2, F_NX, F_NY, F_NZ, F_NXX, F_NYY, F_NZZ, F_NXY_YX, F_NYZ_ZY, F_NZX_XZ,
    F_NXXX, F_NYYY, F_NZZZ, F_SUMPERM, F_2XXY_YXX, F_2XXZ_ZXX, F_2YYZ_ZYY, F_2YYX_XYY, F_2ZZX_XZZ, F_2ZZY_YZZ,
    F_COUNT
};
static_assert(F_COUNT == 23, "23 floats per child lane");

struct BoxData
{
    // SoA over the 4 child lanes, like upstream's v4uf members.
    float f[F_COUNT][BVH_N];
};

void triangle_local_data(V3 a, V3 b, V3 c, LocalData& d)
{
    for (int k = 0; k < 3; ++k) {
        d.box.lo[k] = std::min(a[k], std::min(b[k], c[k]));
        d.box.hi[k] = std::max(a[k], std::max(b[k], c[k]));
    }
    const V3 ab = b - a, ac = c - a;
    const V3 N = 0.5f * cross(ab, ac);
    const float area2 = length2(N);
    const float area = std::sqrt(area2);
    const V3 P = (a + b + c) / 3.0f;
    d.avgP = P;
    d.areaP = P * area;
    d.N = N;
    d.area = area;
    // A triangle has zero first-order tensor about its own centroid.
    d.NijDiag = {0, 0, 0};
    d.Nxy = d.Nyx = d.Nyz = d.Nzy = d.Nzx = d.Nxz = 0;
    d.NijkDiag = {0, 0, 0};
    d.SumPermuteNxyz = 0;
    d.N2xxy_yxx = d.N2xxz_zxx = d.N2yyz_zyy = d.N2yyx_xyy = d.N2zzx_xzz = d.N2zzy_yzz = 0;
    if (area == 0) return;
    const V3 n = N / area;
    // integral over the triangle of (y-P)_j (y-P)_k dA = area/12 * sum over vertices of (v-P)_j (v-P)_k
    // (+: upstream evaluates the same integrals by an axis-sorted split; closed form is mathematically identical).
    const V3 da = a - P, db = b - P, dc = c - P;
    const float s = area / 12.0f;
    const float ixx = s * (da.x * da.x + db.x * db.x + dc.x * dc.x);
    const float iyy = s * (da.y * da.y + db.y * db.y + dc.y * dc.y);
    const float izz = s * (da.z * da.z + db.z * db.z + dc.z * dc.z);
    const float ixy = s * (da.x * da.y + db.x * db.y + dc.x * dc.y);
    const float iyz = s * (da.y * da.z + db.y * db.z + dc.y * dc.z);
    const float izx = s * (da.z * da.x + db.z * db.x + dc.z * dc.x);
    d.NijkDiag = {n.x * ixx, n.y * iyy, n.z * izz};
    d.SumPermuteNxyz = 2.0f * (n.x * iyz + n.y * izx + n.z * ixy);
    d.N2xxy_yxx = 2.0f * (n.x * ixy) + n.y * ixx;
    d.N2xxz_zxx = 2.0f * (n.x * izx) + n.z * ixx;
    d.N2yyz_zyy = 2.0f * (n.y * iyz) + n.z * iyy;
    d.N2yyx_xyy = 2.0f * (n.y * ixy) + n.x * iyy;
    d.N2zzx_xzz = 2.0f * (n.z * izx) + n.x * izz;
    d.N2zzy_yzz = 2.0f * (n.z * iyz) + n.y * izz;
}

// Fill one child lane of the parent's BoxData from the child's LocalData (A.4).
void store_lane(BoxData& bd, int lane, const LocalData& c)
{
    float* f[F_COUNT];
    for (int k = 0; k < F_COUNT; ++k) f[k] = &bd.f[k][lane];
    *f[F_PX] = c.avgP.x;
    *f[F_PY] = c.avgP.y;
    *f[F_PZ] = c.avgP.z;
    V3 lo{c.box.lo[0], c.box.lo[1], c.box.lo[2]}, hi{c.box.hi[0], c.box.hi[1], c.box.hi[2]};
    const V3 maxPDiff = vmax(c.avgP - lo, hi - c.avgP);
    *f[F_R2] = length2(maxPDiff);
    *f[F_NX] = c.N.x;
    *f[F_NY] = c.N.y;
    *f[F_NZ] = c.N.z;
    *f[F_NXX] = c.NijDiag.x;
    *f[F_NYY] = c.NijDiag.y;
    *f[F_NZZ] = c.NijDiag.z;
    *f[F_NXY_YX] = c.Nxy + c.Nyx;
    *f[F_NYZ_ZY] = c.Nyz + c.Nzy;
    *f[F_NZX_XZ] = c.Nzx + c.Nxz;
    *f[F_NXXX] = c.NijkDiag.x;
    *f[F_NYYY] = c.NijkDiag.y;
    *f[F_NZZZ] = c.NijkDiag.z;
    *f[F_SUMPERM] = c.SumPermuteNxyz;
    *f[F_2XXY_YXX] = c.N2xxy_yxx;
    *f[F_2XXZ_ZXX] = c.N2xxz_zxx;
    *f[F_2YYZ_ZYY] = c.N2yyz_zyy;
    *f[F_2YYX_XYY] = c.N2yyx_xyy;
    *f[F_2ZZX_XZZ] = c.N2zzx_xzz;
    *f[F_2ZZY_YZZ] = c.N2zzy_yzz;
}

// Merge nchildren LocalData into the parent's (A.3). The op order below is mirrored by the CUDA moment pass
// (lagrange_b200/csrc/wn_device.cuh, wn_merge_children) so that both produce the same floats on the same topology.
void merge_children(const LocalData* ch, int nchildren, LocalData& out)
{
    V3 N = ch[0].N;
    V3 areaP = ch[0].areaP;
    float area = ch[0].area;
    Box box = ch[0].box;
    for (int i = 1; i < nchildren; ++i) {
        N = N + ch[i].N;
        areaP = areaP + ch[i].areaP;
        area += ch[i].area;
        box.combine(ch[i].box);
    }
    out.N = N;
    out.areaP = areaP;
    out.area = area;
    out.box = box;
    V3 avgP;
    if (area > 0)
        avgP = areaP / area;
    else
        avgP = 0.5f * (V3{box.lo[0], box.lo[1], box.lo[2]} + V3{box.hi[0], box.hi[1], box.hi[2]});
    out.avgP = avgP;

    V3 NijDiag{0, 0, 0};
    float Nxy = 0, Nyx = 0, Nyz = 0, Nzy = 0, Nzx = 0, Nxz = 0;
    V3 NijkDiag{0, 0, 0};
    float S = 0, Bxy = 0, Bxz = 0, Byz = 0, Byx = 0, Bzx = 0, Bzy = 0;
    for (int i = 0; i < nchildren; ++i) {
        const LocalData& c = ch[i];
        const V3 d = c.avgP - avgP;
        const V3 n = c.N;
        // first order: N_ij' = N_ij + N_i d_j
        NijDiag = NijDiag + (c.NijDiag + n * d);
        Nxy += c.Nxy + n.x * d.y;
        Nyx += c.Nyx + n.y * d.x;
        Nyz += c.Nyz + n.y * d.z;
        Nzy += c.Nzy + n.z * d.y;
        Nzx += c.Nzx + n.z * d.x;
        Nxz += c.Nxz + n.x * d.z;
        // second order: N_ijk' = N_ijk + N_ij d_k + N_ik d_j + N_i d_j d_k  (child's un-shifted N_ij)
        NijkDiag = NijkDiag + (c.NijkDiag + 2.0f * (d * c.NijDiag) + (d * d) * n);
        S += c.SumPermuteNxyz +
             2.0f * (d.z * (c.Nxy + c.Nyx) + d.y * (c.Nxz + c.Nzx) + d.x * (c.Nyz + c.Nzy) + n.x * d.y * d.z +
                     n.y * d.z * d.x + n.z * d.x * d.y);
        // B_ij = 2 N_iij + N_jii :  += 2 N_ii d_j + 2 (N_ij + N_ji) d_i + 2 N_i d_i d_j + N_j d_i^2
        Bxy += c.N2xxy_yxx + 2.0f * c.NijDiag.x * d.y + 2.0f * (c.Nxy + c.Nyx) * d.x + 2.0f * n.x * d.x * d.y + n.y * d.x * d.x;
        Bxz += c.N2xxz_zxx + 2.0f * c.NijDiag.x * d.z + 2.0f * (c.Nxz + c.Nzx) * d.x + 2.0f * n.x * d.x * d.z + n.z * d.x * d.x;
        Byz += c.N2yyz_zyy + 2.0f * c.NijDiag.y * d.z + 2.0f * (c.Nyz + c.Nzy) * d.y + 2.0f * n.y * d.y * d.z + n.z * d.y * d.y;
        Byx += c.N2yyx_xyy + 2.0f * c.NijDiag.y * d.x + 2.0f * (c.Nyx + c.Nxy) * d.y + 2.0f * n.y * d.y * d.x + n.x * d.y * d.y;
        Bzx += c.N2zzx_xzz + 2.0f * c.NijDiag.z * d.x + 2.0f * (c.Nzx + c.Nxz) * d.z + 2.0f * n.z * d.z * d.x + n.x * d.z * d.z;
        Bzy += c.N2zzy_yzz + 2.0f * c.NijDiag.z * d.y + 2.0f * (c.Nzy + c.Nyz) * d.z + 2.0f * n.z * d.z * d.y + n.y * d.z * d.z;
    }
    out.NijDiag = NijDiag;
    out.Nxy = Nxy;
    out.Nyx = Nyx;
    out.Nyz = Nyz;
    out.Nzy = Nzy;
    out.Nzx = Nzx;
    out.Nxz = Nxz;
    out.NijkDiag = NijkDiag;
    out.SumPermuteNxyz = S;
    out.N2xxy_yxx = Bxy;
    out.N2xxz_zxx = Bxz;
    out.N2yyz_zyy = Byz;
    out.N2yyx_xyy = Byx;
    out.N2zzx_xzz = Bzx;
    out.N2zzy_yzz = Bzy;
}

struct Counters
{
    uint64_t tests = 0, approx = 0, exact = 0;
};

// UTsignedSolidAngleTri (A.1), float.
inline float tri_solid_angle_f(V3 a, V3 b, V3 c, V3 q)
{
    V3 qa = a - q, qb = b - q, qc = c - q;
    const float al = std::sqrt(length2(qa)), bl = std::sqrt(length2(qb)), cl = std::sqrt(length2(qc));
    if (al == 0 || bl == 0 || cl == 0) return 0.0f;
    qa = qa / al;
    qb = qb / bl;
    qc = qc / cl;
    const float num = dot(qa, cross(qb - qa, qc - qa));
    if (num == 0) return 0.0f;
    const float den = 1.0f + dot(qa, qb) + dot(qa, qc) + dot(qb, qc);
    return 2.0f * std::atan2(num, den);
}

inline double tri_solid_angle_d(const double a[3], const double b[3], const double c[3], const double q[3])
{
    double qa[3], qb[3], qc[3];
    for (int k = 0; k < 3; ++k) {
        qa[k] = a[k] - q[k];
        qb[k] = b[k] - q[k];
        qc[k] = c[k] - q[k];
    }
    const double al = std::sqrt(qa[0] * qa[0] + qa[1] * qa[1] + qa[2] * qa[2]);
    const double bl = std::sqrt(qb[0] * qb[0] + qb[1] * qb[1] + qb[2] * qb[2]);
    const double cl = std::sqrt(qc[0] * qc[0] + qc[1] * qc[1] + qc[2] * qc[2]);
    if (al == 0 || bl == 0 || cl == 0) return 0.0;
    for (int k = 0; k < 3; ++k) {
        qa[k] /= al;
        qb[k] /= bl;
        qc[k] /= cl;
    }
    const double u[3] = {qb[0] - qa[0], qb[1] - qa[1], qb[2] - qa[2]};
    const double v[3] = {qc[0] - qa[0], qc[1] - qa[1], qc[2] - qa[2]};
    const double cr[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double num = qa[0] * cr[0] + qa[1] * cr[1] + qa[2] * cr[2];
    if (num == 0) return 0.0;
    const double den = 1.0 + (qa[0] * qb[0] + qa[1] * qb[1] + qa[2] * qb[2]) + (qa[0] * qc[0] + qa[1] * qc[1] + qa[2] * qc[2]) +
                       (qb[0] * qc[0] + qb[1] * qc[1] + qb[2] * qc[2]);
    return 2.0 * std::atan2(num, den);
}

struct RefEngine
{
    std::vector<V3> verts;
    std::vector<std::array<int32_t, 3>> tris;
    std::vector<Node> nodes;
    std::vector<BoxData> data;
    int order = 2;
    double build_seconds = 0;
    double bvh_seconds = 0;

    // post-order moment pass (UT_SolidAngle::init's traverse functors restated)
    void moments(int nodei, LocalData& out)
    {
        LocalData ch[BVH_N];
        int n = 0;
        for (int s = 0; s < BVH_N; ++s) {
            const int32_t c = nodes[nodei].child[s];
            if (c == kEmpty) break; // empty slots are trailing
            if (is_tri(c)) {
                const auto& t = tris[dec_tri(c)];
                triangle_local_data(verts[t[0]], verts[t[1]], verts[t[2]], ch[n]);
            } else {
                moments(c, ch[n]);
            }
            ++n;
        }
        BoxData& bd = data[nodei];
        std::memset(&bd, 0, sizeof(bd));
        for (int s = 0; s < n; ++s) store_lane(bd, s, ch[s]);
        // non-existent children: infinite radius => never approximated, traversal sees EMPTY
        for (int s = n; s < BVH_N; ++s) bd.f[F_R2][s] = std::numeric_limits<float>::infinity();
        if (order < 2) {
            for (int k = F_NXXX; k < F_COUNT; ++k)
                for (int s = 0; s < BVH_N; ++s) bd.f[k][s] = 0;
        }
        if (order < 1) {
            for (int k = F_NXX; k < F_NXXX; ++k)
                for (int s = 0; s < BVH_N; ++s) bd.f[k][s] = 0;
        }
        merge_children(ch, n, out);
    }

    void build()
    {
        const auto t0 = std::chrono::steady_clock::now();
        const int n = int(tris.size());
        nodes.clear();
        data.clear();
        if (n == 0) return;
        std::vector<Box> boxes(n);
        Box all;
        all.init_empty();
        for (int i = 0; i < n; ++i) {
            Box b;
            b.init_empty();
            for (int k = 0; k < 3; ++k) {
                const V3 p = verts[tris[i][k]];
                for (int a = 0; a < 3; ++a) {
                    b.lo[a] = std::min(b.lo[a], p[a]);
                    b.hi[a] = std::max(b.hi[a], p[a]);
                }
            }
            boxes[i] = b;
            all.combine(b);
        }
        std::vector<int32_t> indices(n);
        std::iota(indices.begin(), indices.end(), 0);
        nodes.reserve(size_t(n) / 2 + 16);
        nodes.push_back(Node{{kEmpty, kEmpty, kEmpty, kEmpty}});
        BvhBuilder b{boxes.data(), &nodes};
        b.init_node(0, all, indices.data(), n);
        const auto t1 = std::chrono::steady_clock::now();
        bvh_seconds = std::chrono::duration<double>(t1 - t0).count();
        data.resize(nodes.size());
        LocalData root;
        moments(0, root);
        const auto t2 = std::chrono::steady_clock::now();
        build_seconds = std::chrono::duration<double>(t2 - t0).count();
    }

    // UT_SolidAngle::computeSolidAngle restated (A.5): depth first, 4 child lanes per node, float.
    float solid_angle_node(int nodei, V3 q, float beta2, Counters* cnt) const
    {
        const BoxData& d = data[nodei];
        const Node& nd = nodes[nodei];
        float omega_lane[BVH_N];
        unsigned descend = 0;
        int nlanes = 0;
#if defined(__SSE2__)
        if (g_simd_lanes) {
            // all four lanes at once; empty lanes carry zeros and an infinite radius (=> "descend", ignored below)
            typedef float v4f __attribute__((vector_size(16)));
            auto LD = [&](int k) {
                v4f v;
                std::memcpy(&v, d.f[k], sizeof(v));
                return v;
            };
            const v4f qx = {q.x, q.x, q.x, q.x}, qy = {q.y, q.y, q.y, q.y}, qz = {q.z, q.z, q.z, q.z};
            v4f rx = qx - LD(F_PX), ry = qy - LD(F_PY), rz = qz - LD(F_PZ);
            const v4f ql2 = rx * rx + ry * ry + rz * rz;
            const v4f thr = LD(F_R2) * beta2;
            const v4f one = {1.0f, 1.0f, 1.0f, 1.0f};
            const v4f m2 = one / ql2;
            const v4f m1 = (v4f)_mm_sqrt_ps((__m128)m2);
            rx = rx * m1;
            ry = ry * m1;
            rz = rz * m1;
            v4f om = -m2 * (rx * LD(F_NX) + ry * LD(F_NY) + rz * LD(F_NZ));
            if (order >= 1) {
                const v4f q2x = rx * rx, q2y = ry * ry, q2z = rz * rz;
                const v4f m3 = m2 * m1;
                const v4f nxx = LD(F_NXX), nyy = LD(F_NYY), nzz = LD(F_NZZ);
                const v4f o1 = m3 * (nxx + nyy + nzz -
                                     3.0f * ((q2x * nxx + q2y * nyy + q2z * nzz) + rx * ry * LD(F_NXY_YX) + rx * rz * LD(F_NZX_XZ) +
                                             ry * rz * LD(F_NYZ_ZY)));
                om += o1;
                if (order >= 2) {
                    const v4f q3x = q2x * rx, q3y = q2y * ry, q3z = q2z * rz;
                    const v4f m4 = m2 * m2;
                    const v4f bxxy = LD(F_2XXY_YXX), bxxz = LD(F_2XXZ_ZXX), byyz = LD(F_2YYZ_ZYY), byyx = LD(F_2YYX_XYY), bzzx = LD(F_2ZZX_XZZ),
                              bzzy = LD(F_2ZZY_YZZ);
                    const v4f t0x = byyx + bzzx, t0y = bzzy + bxxy, t0z = bxxz + byyz;
                    const v4f t1x = ry * bxxy + rz * bxxz, t1y = rz * byyz + rx * byyx, t1z = rx * bzzx + ry * bzzy;
                    const v4f dgx = LD(F_NXXX), dgy = LD(F_NYYY), dgz = LD(F_NZZZ);
                    const v4f o2 = m4 * (1.5f * (rx * (3.0f * dgx + t0x) + ry * (3.0f * dgy + t0y) + rz * (3.0f * dgz + t0z)) -
                                         7.5f * ((q3x * dgx + q3y * dgy + q3z * dgz) + rx * ry * rz * LD(F_SUMPERM) +
                                                 (q2x * t1x + q2y * t1y + q2z * t1z)));
                    om += o2;
                }
            }
            for (int s = 0; s < BVH_N; ++s) {
                if (nd.child[s] == kEmpty) {
                    omega_lane[s] = 0;
                    continue;
                }
                ++nlanes;
                const bool desc = ql2[s] <= thr[s];
                if (cnt) ++cnt->tests;
                if (!desc && std::isfinite(om[s])) {
                    omega_lane[s] = om[s];
                    if (cnt) ++cnt->approx;
                } else {
                    omega_lane[s] = 0;
                    descend |= 1u << s;
                }
            }
        } else
#endif
        {
        for (int s = 0; s < BVH_N; ++s) {
            if (nd.child[s] == kEmpty) {
                omega_lane[s] = 0;
                continue;
            }
            ++nlanes;
            V3 r{q.x - d.f[F_PX][s], q.y - d.f[F_PY][s], q.z - d.f[F_PZ][s]};
            const float ql2 = r.x * r.x + r.y * r.y + r.z * r.z;
            const bool desc = ql2 <= d.f[F_R2][s] * beta2;
            float om = 0;
            if (!desc) {
                const float m2 = 1.0f / ql2;
                const float m1 = std::sqrt(m2);
                r = r * m1;
                om = -m2 * (r.x * d.f[F_NX][s] + r.y * d.f[F_NY][s] + r.z * d.f[F_NZ][s]);
                if (order >= 1) {
                    const V3 q2 = r * r;
                    const float m3 = m2 * m1;
                    const float o1 =
                        m3 * (d.f[F_NXX][s] + d.f[F_NYY][s] + d.f[F_NZZ][s] -
                              3.0f * ((q2.x * d.f[F_NXX][s] + q2.y * d.f[F_NYY][s] + q2.z * d.f[F_NZZ][s]) +
                                      r.x * r.y * d.f[F_NXY_YX][s] + r.x * r.z * d.f[F_NZX_XZ][s] + r.y * r.z * d.f[F_NYZ_ZY][s]));
                    om += o1;
                    if (order >= 2) {
                        const V3 q3 = q2 * r;
                        const float m4 = m2 * m2;
                        const V3 t0{d.f[F_2YYX_XYY][s] + d.f[F_2ZZX_XZZ][s], d.f[F_2ZZY_YZZ][s] + d.f[F_2XXY_YXX][s],
                                    d.f[F_2XXZ_ZXX][s] + d.f[F_2YYZ_ZYY][s]};
                        const V3 t1{r.y * d.f[F_2XXY_YXX][s] + r.z * d.f[F_2XXZ_ZXX][s],
                                    r.z * d.f[F_2YYZ_ZYY][s] + r.x * d.f[F_2YYX_XYY][s],
                                    r.x * d.f[F_2ZZX_XZZ][s] + r.y * d.f[F_2ZZY_YZZ][s]};
                        const V3 diag{d.f[F_NXXX][s], d.f[F_NYYY][s], d.f[F_NZZZ][s]};
                        const float o2 = m4 * (1.5f * dot(r, 3.0f * diag + t0) -
                                               7.5f * (dot(q3, diag) + r.x * r.y * r.z * d.f[F_SUMPERM][s] + dot(q2, t1)));
                        om += o2;
                    }
                }
            }
            if (cnt) ++cnt->tests;
            // non-finite approximations are discarded and the lane descends instead
            if (!desc && std::isfinite(om)) {
                omega_lane[s] = om;
                if (cnt) ++cnt->approx;
            } else {
                omega_lane[s] = 0;
                descend |= 1u << s;
            }
        }
        }
        float sum = omega_lane[0];
        for (int s = 1; s < BVH_N; ++s) sum += omega_lane[s];
        if (!descend) return sum;
        float child_sum[BVH_N] = {0, 0, 0, 0};
        for (int s = 0; s < nlanes; ++s) {
            child_sum[s] = 0;
            if (!((descend >> s) & 1)) continue;
            const int32_t c = nd.child[s];
            if (is_tri(c)) {
                const auto& t = tris[dec_tri(c)];
                child_sum[s] = tri_solid_angle_f(verts[t[0]], verts[t[1]], verts[t[2]], q);
                if (cnt) ++cnt->exact;
            } else {
                child_sum[s] = solid_angle_node(c, q, beta2, cnt);
            }
        }
        float post = (descend & 1) ? child_sum[0] : 0.0f;
        for (int s = 1; s < nlanes; ++s) post += ((descend >> s) & 1) ? child_sum[s] : 0.0f;
        return sum + post;
    }

    float solid_angle(V3 q, float beta, Counters* cnt) const
    {
        if (nodes.empty()) return 0.0f;
        return solid_angle_node(0, q, beta * beta, cnt);
    }
};

// modules/winding/src/FastWindingNumber.cpp:66 :  computeSolidAngle(q) / (4.f * lagrange::internal::pi) > 0.5f
// with pi a constexpr double (modules/core/include/lagrange/internal/constants.h:16): evaluated in double.
inline bool inside_predicate(float omega)
{
    constexpr double pi = 3.14159265358979323846264338327950288;
    return omega / (4.f * pi) > 0.5f;
}

int resolve_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    return nthreads;
}

} // namespace

extern "C" {

void wno_set_simd_lanes(int on)
{
    g_simd_lanes = on ? 1 : 0;
}

int wno_num_threads()
{
    return resolve_threads(0);
}

// ---- exact64 ------------------------------------------------------------------------------------------------
int wno_exact64(const float* v, int64_t nV, const int32_t* tri, int64_t nT, const float* q, int64_t nQ, double* out, int nthreads)
{
    (void)nV;
    nthreads = resolve_threads(nthreads);
    std::vector<double> tv(size_t(nT) * 9);
    for (int64_t t = 0; t < nT; ++t)
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) tv[size_t(t) * 9 + k * 3 + a] = double(v[size_t(tri[t * 3 + k]) * 3 + a]);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t i = 0; i < nQ; ++i) {
        const double qq[3] = {double(q[i * 3]), double(q[i * 3 + 1]), double(q[i * 3 + 2])};
        // Kahan-free pairwise-ish: accumulate in long double to keep the sum honest for 1e6+ triangles
        long double acc = 0;
        for (int64_t t = 0; t < nT; ++t) acc += tri_solid_angle_d(&tv[t * 9], &tv[t * 9 + 3], &tv[t * 9 + 6], qq);
        out[i] = double(acc);
    }
    return 0;
}

// ---- unsigned distance (ground truth for the narrow-band distance kernel K10) -------------
// Double precision, brute force over all triangles. Deliberately a different formulation from the kernel's region walk:
// the foot of the perpendicular if it falls inside the triangle, else the nearest of the three edges.
static double seg_dist2_d(const double* p, const double* a, const double* b)
{
    const double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    const double ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
    const double len2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
    double t = len2 > 0 ? (ap[0] * ab[0] + ap[1] * ab[1] + ap[2] * ab[2]) / len2 : 0.0;
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    const double d[3] = {ap[0] - t * ab[0], ap[1] - t * ab[1], ap[2] - t * ab[2]};
    return d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
}

static double tri_dist2_d(const double* a, const double* b, const double* c, const double* p)
{
    const double e0[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    const double e1[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const double n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
    const double n2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    if (n2 > 0) {
        const double ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
        // barycentric coordinates of the projection: u = n . (e0 x ap)... via the two cross products with the normal
        const double c0[3] = {ap[1] * e1[2] - ap[2] * e1[1], ap[2] * e1[0] - ap[0] * e1[2], ap[0] * e1[1] - ap[1] * e1[0]};
        const double c1[3] = {e0[1] * ap[2] - e0[2] * ap[1], e0[2] * ap[0] - e0[0] * ap[2], e0[0] * ap[1] - e0[1] * ap[0]};
        const double v = (c0[0] * n[0] + c0[1] * n[1] + c0[2] * n[2]) / n2; // weight of b
        const double w = (c1[0] * n[0] + c1[1] * n[1] + c1[2] * n[2]) / n2; // weight of c
        if (v >= 0 && w >= 0 && v + w <= 1) {
            const double h = ap[0] * n[0] + ap[1] * n[1] + ap[2] * n[2];
            return h * h / n2;
        }
    }
    return std::min(seg_dist2_d(p, a, b), std::min(seg_dist2_d(p, b, c), seg_dist2_d(p, c, a)));
}

int wno_distance64(const float* v, int64_t nV, const int32_t* tri, int64_t nT, const float* q, int64_t nQ, double* out, int nthreads)
{
    (void)nV;
    nthreads = resolve_threads(nthreads);
    std::vector<double> tv(size_t(nT) * 9);
    for (int64_t t = 0; t < nT; ++t)
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) tv[size_t(t) * 9 + k * 3 + a] = double(v[size_t(tri[t * 3 + k]) * 3 + a]);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t i = 0; i < nQ; ++i) {
        const double qq[3] = {double(q[i * 3]), double(q[i * 3 + 1]), double(q[i * 3 + 2])};
        double best = std::numeric_limits<double>::infinity();
        for (int64_t t = 0; t < nT; ++t) best = std::min(best, tri_dist2_d(&tv[t * 9], &tv[t * 9 + 3], &tv[t * 9 + 6], qq));
        out[i] = std::sqrt(best);
    }
    return 0;
}

// ---- reference restatement ------------------------------------------------------------------------------------
void* wno_ref_create(const float* v, int64_t nV, const int32_t* tri, int64_t nT, int order)
{
    auto* e = new RefEngine;
    e->order = order;
    e->verts.resize(size_t(nV));
    for (int64_t i = 0; i < nV; ++i) e->verts[i] = {v[i * 3], v[i * 3 + 1], v[i * 3 + 2]};
    e->tris.resize(size_t(nT));
    for (int64_t i = 0; i < nT; ++i) e->tris[i] = {tri[i * 3], tri[i * 3 + 1], tri[i * 3 + 2]};
    e->build();
    return e;
}

void wno_ref_destroy(void* h)
{
    delete static_cast<RefEngine*>(h);
}

int64_t wno_ref_num_nodes(void* h)
{
    return int64_t(static_cast<RefEngine*>(h)->nodes.size());
}

double wno_ref_build_seconds(void* h)
{
    return static_cast<RefEngine*>(h)->build_seconds;
}

// child[n_nodes * 4], neutral encoding (>=0 internal, -1 empty, <=-2 triangle -(c+2))
int wno_ref_get_topology(void* h, int32_t* child)
{
    auto* e = static_cast<RefEngine*>(h);
    for (size_t i = 0; i < e->nodes.size(); ++i)
        for (int s = 0; s < BVH_N; ++s) child[i * BVH_N + s] = e->nodes[i].child[s];
    return 0;
}

// out[n_nodes * 4 * 23]: for node i, slot s, the 23 stored floats of that child lane (A.4 order, enum above)
int wno_ref_get_boxdata(void* h, float* out)
{
    auto* e = static_cast<RefEngine*>(h);
    for (size_t i = 0; i < e->nodes.size(); ++i)
        for (int s = 0; s < BVH_N; ++s)
            for (int k = 0; k < F_COUNT; ++k) out[(i * BVH_N + s) * F_COUNT + k] = e->data[i].f[k][s];
    return 0;
}

// counters (optional, may be null): totals over all queries of {lanes tested, lanes approximated, exact triangles}
int wno_ref_solid_angle(void* h, const float* q, int64_t nQ, float beta, float* out, uint64_t* counters, int nthreads)
{
    auto* e = static_cast<RefEngine*>(h);
    nthreads = resolve_threads(nthreads);
    uint64_t T = 0, A = 0, E = 0;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) reduction(+ : T, A, E)
    for (int64_t i = 0; i < nQ; ++i) {
        Counters c;
        out[i] = e->solid_angle({q[i * 3], q[i * 3 + 1], q[i * 3 + 2]}, beta, counters ? &c : nullptr);
        T += c.tests;
        A += c.approx;
        E += c.exact;
    }
    if (counters) {
        counters[0] = T;
        counters[1] = A;
        counters[2] = E;
    }
    return 0;
}

int wno_ref_is_inside(void* h, const float* q, int64_t nQ, float beta, uint8_t* out, int nthreads)
{
    auto* e = static_cast<RefEngine*>(h);
    nthreads = resolve_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (int64_t i = 0; i < nQ; ++i)
        out[i] = inside_predicate(e->solid_angle({q[i * 3], q[i * 3 + 1], q[i * 3 + 2]}, beta, nullptr)) ? 1 : 0;
    return 0;
}

// Implicit cell-centred lattice like mesh_to_volume's transform (modules/volume/src/mesh_to_volume.cpp:147-149):
// p = origin + spacing * (ijk + 0.5), x fastest. Evaluates indices [first, first + count) of the x-fastest order
// with stride `stride` (so a bounded sample of a big lattice can be timed). Writes count results.
int wno_ref_is_inside_grid(void* h, const float* origin, const float* spacing, const int64_t* dims, int64_t first, int64_t count,
                           int64_t stride, float beta, uint8_t* out_inside, float* out_omega, int nthreads)
{
    auto* e = static_cast<RefEngine*>(h);
    nthreads = resolve_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (int64_t k = 0; k < count; ++k) {
        const int64_t idx = first + k * stride;
        const int64_t ix = idx % dims[0], iy = (idx / dims[0]) % dims[1], iz = idx / (dims[0] * dims[1]);
        const V3 p{origin[0] + spacing[0] * (float(ix) + 0.5f), origin[1] + spacing[1] * (float(iy) + 0.5f),
                   origin[2] + spacing[2] * (float(iz) + 0.5f)};
        const float om = e->solid_angle(p, beta, nullptr);
        if (out_omega) out_omega[k] = om;
        if (out_inside) out_inside[k] = inside_predicate(om) ? 1 : 0;
    }
    return 0;
}

// The predicate alone, so tests can pin the float threshold the CUDA path uses against the reference expression.
int wno_inside_predicate(float omega)
{
    return inside_predicate(omega) ? 1 : 0;
}

// float exact brute force with the reference's triangle formula (the tiled CUDA kernel's arithmetic twin)
int wno_exact32(const float* v, int64_t nV, const int32_t* tri, int64_t nT, const float* q, int64_t nQ, float* out, int nthreads)
{
    (void)nV;
    nthreads = resolve_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t i = 0; i < nQ; ++i) {
        const V3 qq{q[i * 3], q[i * 3 + 1], q[i * 3 + 2]};
        float acc = 0;
        for (int64_t t = 0; t < nT; ++t) {
            const int32_t* tt = tri + t * 3;
            const V3 a{v[tt[0] * 3], v[tt[0] * 3 + 1], v[tt[0] * 3 + 2]};
            const V3 b{v[tt[1] * 3], v[tt[1] * 3 + 1], v[tt[1] * 3 + 2]};
            const V3 c{v[tt[2] * 3], v[tt[2] * 3 + 1], v[tt[2] * 3 + 2]};
            acc += tri_solid_angle_f(a, b, c, qq);
        }
        out[i] = acc;
    }
    return 0;
}

} // extern "C"
