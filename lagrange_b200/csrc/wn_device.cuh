// wn_device.cuh — arithmetic shared by every kernel of the engine, written as host/device functions so that the very
// same source is (a) inlined into the sm_100a kernels and (b) compiled by g++ into the host emulation harness that
// tests/ uses to check the build and packing logic on machines without a GPU (tests/emul/wn_emul.cpp).
//
// What each block restates (the upstream UT_SolidAngle source is not in the reference tree):
//   wn_tri_solid_angle      A.1  UTsignedSolidAngleTri, called from the leaves of computeSolidAngle
//                                (modules/winding/src/FastWindingNumber.cpp:66,75)
//   wn_tri_local            A.2  per-triangle moments about its own centroid
//   wn_merge_children       A.3  bottom-up merge with the parallel-axis shift
//   wn_local_to_ref23       A.4  the 23 floats the reference stores per child lane
//   wn_pack_record          A.4 folded: |r^|=1 lets the trace term join the quadratic form and the linear order-2
//                                term join the cubic form, so a record is P,R2 | N | 6 quadratic | 10 cubic = 23 floats
//   wn_eval_record          A.5  far-field evaluation (order <= 2)
//   wn_inside_from_omega    FastWindingNumber.cpp:66  (double)omega / (4.0*pi) > 0.5
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define WN_HD __host__ __device__ __forceinline__
#else
#define WN_HD inline
#endif

// ----------------------------------------------------------------------------------------------------------------
// Types that exist on both sides
// ----------------------------------------------------------------------------------------------------------------
#if !defined(__CUDACC__)
struct alignas(16) float4
{
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w)
{
    return float4{x, y, z, w};
}
struct alignas(16) int4
{
    int x, y, z, w;
};
static inline int4 make_int4(int x, int y, int z, int w)
{
    return int4{x, y, z, w};
}
#endif

struct WnV3
{
    float x, y, z;
};
WN_HD WnV3 wn_v3(float x, float y, float z)
{
    WnV3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}

// The build kernels must round like the CPU oracle (which is compiled with -ffp-contract=off), so that moments on
// the same topology agree bit for bit: on the device every product/sum below goes through __fmul_rn/__fadd_rn, which
// nvcc never contracts into an FMA. The query kernels do not use these (they want FMAs).
#if defined(__CUDA_ARCH__)
#define WN_MUL(a, b) __fmul_rn((a), (b))
#define WN_ADD(a, b) __fadd_rn((a), (b))
#define WN_SUB(a, b) __fsub_rn((a), (b))
#define WN_DIV(a, b) __fdiv_rn((a), (b))
#define WN_SQRT(a) __fsqrt_rn((a))
#define WN_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define WN_MUL(a, b) ((a) * (b))
#define WN_ADD(a, b) ((a) + (b))
#define WN_SUB(a, b) ((a) - (b))
#define WN_DIV(a, b) ((a) / (b))
#define WN_SQRT(a) sqrtf((a))
#define WN_FMA(a, b, c) fmaf((a), (b), (c))
#endif

WN_HD float wn_min(float a, float b)
{
    return a < b ? a : b;
}
WN_HD float wn_max(float a, float b)
{
    return a > b ? a : b;
}

WN_HD int wn_float_as_int(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    union
    {
        float f;
        int i;
    } u;
    u.f = f;
    return u.i;
#endif
}
WN_HD float wn_int_as_float(int i)
{
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    union
    {
        float f;
        int i;
    } u;
    u.i = i;
    return u.f;
#endif
}

// ----------------------------------------------------------------------------------------------------------------
// Hierarchy encoding
// ----------------------------------------------------------------------------------------------------------------
// Child slot encoding at the C-ABI (wn_create_from_topology) and in the device topology tables.
#define WN_CHILD_EMPTY (-1)
WN_HD int wn_enc_tri(int t)
{
    return -(t + 2);
}
WN_HD bool wn_is_tri(int c)
{
    return c <= -2;
}
WN_HD int wn_dec_tri(int c)
{
    return -(c + 2);
}

// Packed depth-first record array ("entries"). For entry i:
//   rec[0][i] = (Px, Py, Pz, R2)          R2's sign bit set <=> leaf entry (R2 itself is >= 0; may be +inf)
//   rec[1][i] = (Nx, Ny, Nz, link[i] as raw bits)
//   rec[2][i] = (qxx, qyy, qzz, qxy)      quadratic form  a1(r^) = sum q_ab r^_a r^_b
//   rec[3][i] = (qyz, qzx, cxxx, cyyy)    cubic form      a2(r^) = sum c_abc r^_a r^_b r^_c
//   rec[4][i] = (czzz, cxyz, cxxy, cxxz)
//   rec[5][i] = (cyyz, cyyx, czzx, czzy)
//   link[i]   = internal: index of the first entry after this subtree (skip link, > i)
//               leaf:     (first_triangle << 4) | (num_triangles - 1)   into the depth-first ordered triangle array
#define WN_LEAF_COUNT_BITS 4
#define WN_MAX_LEAF_SIZE 16
#define WN_MAX_TRIANGLES (1 << 27)

WN_HD int wn_leaf_link(int first, int count)
{
    return (first << WN_LEAF_COUNT_BITS) | (count - 1);
}

// ----------------------------------------------------------------------------------------------------------------
// Morton codes
// ----------------------------------------------------------------------------------------------------------------
WN_HD uint64_t wn_expand_bits21(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
WN_HD uint32_t wn_expand_bits10(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
// One coordinate of a triangle centroid / its position in the unit cube of the scene bounds. Unfused on purpose: nvcc
// would otherwise contract (sum * 1/3) - lo into one FMA and the device's Morton order would differ from the host
// emulation's in the last bit.
WN_HD float wn_centroid_coord(float a, float b, float c)
{
    return WN_MUL(WN_ADD(WN_ADD(a, b), c), 1.0f / 3.0f);
}
WN_HD float wn_unit_coord(float c, float lo, float inv_extent)
{
    return WN_MUL(WN_SUB(c, lo), inv_extent);
}

// p is already normalised to [0,1)^3 (clamped here). bits per axis = 21 (63-bit code) or 10 (30-bit code).
WN_HD uint64_t wn_morton(float px, float py, float pz, int bits_per_axis)
{
    const float scale = bits_per_axis == 21 ? 2097152.0f : 1024.0f;
    const float hi = scale - 1.0f;
    float fx = wn_min(wn_max(WN_MUL(px, scale), 0.0f), hi);
    float fy = wn_min(wn_max(WN_MUL(py, scale), 0.0f), hi);
    float fz = wn_min(wn_max(WN_MUL(pz, scale), 0.0f), hi);
    // NaN -> 0 (comparisons above are false for NaN: wn_max returns b = 0 when a is NaN)
    const uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy, iz = (uint32_t)fz;
    if (bits_per_axis == 21) return (wn_expand_bits21(ix) << 2) | (wn_expand_bits21(iy) << 1) | wn_expand_bits21(iz);
    return ((uint64_t)wn_expand_bits10(ix) << 2) | ((uint64_t)wn_expand_bits10(iy) << 1) | (uint64_t)wn_expand_bits10(iz);
}

WN_HD int wn_clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
WN_HD int wn_clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

// Karras 2012 common-prefix length with index tie break; -1 outside [0, n).
WN_HD int wn_lbvh_delta(const uint64_t* keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + wn_clz32((uint32_t)i ^ (uint32_t)j);
    return wn_clz64(a ^ b);
}

// Build internal node i of the LBVH over n >= 2 sorted keys. Node ids: internal 0..n-2, leaf p -> (n-1) + p.
// Writes child[2i], child[2i+1] (node ids) and the parent/slot of both children.
WN_HD void wn_lbvh_node(const uint64_t* keys, int n, int i, int* child, int* parent, unsigned char* slot)
{
    const int d = (wn_lbvh_delta(keys, n, i, i + 1) - wn_lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = wn_lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (wn_lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (wn_lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = wn_lbvh_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (wn_lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int nI = n - 1;
    const int left = (lo == gamma) ? nI + gamma : gamma;
    const int right = (hi == gamma + 1) ? nI + gamma + 1 : gamma + 1;
    child[2 * i] = left;
    child[2 * i + 1] = right;
    parent[left] = i;
    parent[right] = i;
    slot[left] = 0;
    slot[right] = 1;
}

// ----------------------------------------------------------------------------------------------------------------
// Moments (float32, unfused, same operation order as oracle/wn_oracle.cpp)
// ----------------------------------------------------------------------------------------------------------------
// 36 floats (9 float4) so a record moves as aligned 16-byte pieces.
struct WnLocal
{
    float lo[3], hi[3];           // 0..5   AABB
    float P[3];                   // 6..8   area-weighted centroid
    float areaP[3];               // 9..11
    float N[3];                   // 12..14 area-weighted normal sum
    float area;                   // 15
    float Nd[3];                  // 16..18 Nxx, Nyy, Nzz
    float Nxy, Nyx, Nyz, Nzy, Nzx, Nxz; // 19..24
    float Td[3];                  // 25..27 Nxxx, Nyyy, Nzzz
    float S;                      // 28     2 (Nxyz + Nyzx + Nzxy)
    float Bxy, Bxz, Byz, Byx, Bzx, Bzy; // 29..34  B_ij = 2 N_iij + N_jii
    float pad;                    // 35
};
#define WN_LOCAL_FLOATS 36

WN_HD void wn_tri_local(WnV3 a, WnV3 b, WnV3 c, WnLocal& d)
{
    d.lo[0] = wn_min(a.x, wn_min(b.x, c.x));
    d.lo[1] = wn_min(a.y, wn_min(b.y, c.y));
    d.lo[2] = wn_min(a.z, wn_min(b.z, c.z));
    d.hi[0] = wn_max(a.x, wn_max(b.x, c.x));
    d.hi[1] = wn_max(a.y, wn_max(b.y, c.y));
    d.hi[2] = wn_max(a.z, wn_max(b.z, c.z));
    const float abx = WN_SUB(b.x, a.x), aby = WN_SUB(b.y, a.y), abz = WN_SUB(b.z, a.z);
    const float acx = WN_SUB(c.x, a.x), acy = WN_SUB(c.y, a.y), acz = WN_SUB(c.z, a.z);
    // N = 0.5 * cross(ab, ac)
    const float nx = WN_MUL(0.5f, WN_SUB(WN_MUL(aby, acz), WN_MUL(abz, acy)));
    const float ny = WN_MUL(0.5f, WN_SUB(WN_MUL(abz, acx), WN_MUL(abx, acz)));
    const float nz = WN_MUL(0.5f, WN_SUB(WN_MUL(abx, acy), WN_MUL(aby, acx)));
    const float area2 = WN_ADD(WN_ADD(WN_MUL(nx, nx), WN_MUL(ny, ny)), WN_MUL(nz, nz));
    const float area = WN_SQRT(area2);
    const float px = WN_DIV(WN_ADD(WN_ADD(a.x, b.x), c.x), 3.0f);
    const float py = WN_DIV(WN_ADD(WN_ADD(a.y, b.y), c.y), 3.0f);
    const float pz = WN_DIV(WN_ADD(WN_ADD(a.z, b.z), c.z), 3.0f);
    d.P[0] = px;
    d.P[1] = py;
    d.P[2] = pz;
    d.areaP[0] = WN_MUL(px, area);
    d.areaP[1] = WN_MUL(py, area);
    d.areaP[2] = WN_MUL(pz, area);
    d.N[0] = nx;
    d.N[1] = ny;
    d.N[2] = nz;
    d.area = area;
    d.Nd[0] = d.Nd[1] = d.Nd[2] = 0.0f;
    d.Nxy = d.Nyx = d.Nyz = d.Nzy = d.Nzx = d.Nxz = 0.0f;
    d.Td[0] = d.Td[1] = d.Td[2] = 0.0f;
    d.S = 0.0f;
    d.Bxy = d.Bxz = d.Byz = d.Byx = d.Bzx = d.Bzy = 0.0f;
    d.pad = 0.0f;
    if (area == 0.0f) return;
    const float ux = WN_DIV(nx, area), uy = WN_DIV(ny, area), uz = WN_DIV(nz, area);
    const float dax = WN_SUB(a.x, px), day = WN_SUB(a.y, py), daz = WN_SUB(a.z, pz);
    const float dbx = WN_SUB(b.x, px), dby = WN_SUB(b.y, py), dbz = WN_SUB(b.z, pz);
    const float dcx = WN_SUB(c.x, px), dcy = WN_SUB(c.y, py), dcz = WN_SUB(c.z, pz);
    const float s = WN_DIV(area, 12.0f);
#define WN_SUM3(p, q, r) WN_ADD(WN_ADD((p), (q)), (r))
    const float ixx = WN_MUL(s, WN_SUM3(WN_MUL(dax, dax), WN_MUL(dbx, dbx), WN_MUL(dcx, dcx)));
    const float iyy = WN_MUL(s, WN_SUM3(WN_MUL(day, day), WN_MUL(dby, dby), WN_MUL(dcy, dcy)));
    const float izz = WN_MUL(s, WN_SUM3(WN_MUL(daz, daz), WN_MUL(dbz, dbz), WN_MUL(dcz, dcz)));
    const float ixy = WN_MUL(s, WN_SUM3(WN_MUL(dax, day), WN_MUL(dbx, dby), WN_MUL(dcx, dcy)));
    const float iyz = WN_MUL(s, WN_SUM3(WN_MUL(day, daz), WN_MUL(dby, dbz), WN_MUL(dcy, dcz)));
    const float izx = WN_MUL(s, WN_SUM3(WN_MUL(daz, dax), WN_MUL(dbz, dbx), WN_MUL(dcz, dcx)));
    d.Td[0] = WN_MUL(ux, ixx);
    d.Td[1] = WN_MUL(uy, iyy);
    d.Td[2] = WN_MUL(uz, izz);
    d.S = WN_MUL(2.0f, WN_SUM3(WN_MUL(ux, iyz), WN_MUL(uy, izx), WN_MUL(uz, ixy)));
    d.Bxy = WN_ADD(WN_MUL(2.0f, WN_MUL(ux, ixy)), WN_MUL(uy, ixx));
    d.Bxz = WN_ADD(WN_MUL(2.0f, WN_MUL(ux, izx)), WN_MUL(uz, ixx));
    d.Byz = WN_ADD(WN_MUL(2.0f, WN_MUL(uy, iyz)), WN_MUL(uz, iyy));
    d.Byx = WN_ADD(WN_MUL(2.0f, WN_MUL(uy, ixy)), WN_MUL(ux, iyy));
    d.Bzx = WN_ADD(WN_MUL(2.0f, WN_MUL(uz, izx)), WN_MUL(ux, izz));
    d.Bzy = WN_ADD(WN_MUL(2.0f, WN_MUL(uz, iyz)), WN_MUL(uy, izz));
}

// B_ij' = B_ij + 2 N_ii d_j + 2 (N_ij + N_ji) d_i + 2 N_i d_i d_j + N_j d_i^2   (left-to-right sum like the oracle)
WN_HD float wn_shift_B(float B, float Nii, float Nij, float Nji, float Ni, float Nj, float di, float dj)
{
    float r = WN_ADD(B, WN_MUL(WN_MUL(2.0f, Nii), dj));
    r = WN_ADD(r, WN_MUL(WN_MUL(2.0f, WN_ADD(Nij, Nji)), di));
    r = WN_ADD(r, WN_MUL(WN_MUL(WN_MUL(2.0f, Ni), di), dj));
    r = WN_ADD(r, WN_MUL(WN_MUL(Nj, di), di));
    return r;
}

// Merge n children (slot order) into `out`. Mirrors merge_children() in oracle/wn_oracle.cpp.
WN_HD void wn_merge_children(const WnLocal* ch, int n, WnLocal& out)
{
    float Nx = ch[0].N[0], Ny = ch[0].N[1], Nz = ch[0].N[2];
    float apx = ch[0].areaP[0], apy = ch[0].areaP[1], apz = ch[0].areaP[2];
    float area = ch[0].area;
    float lo0 = ch[0].lo[0], lo1 = ch[0].lo[1], lo2 = ch[0].lo[2];
    float hi0 = ch[0].hi[0], hi1 = ch[0].hi[1], hi2 = ch[0].hi[2];
    for (int i = 1; i < n; ++i) {
        Nx = WN_ADD(Nx, ch[i].N[0]);
        Ny = WN_ADD(Ny, ch[i].N[1]);
        Nz = WN_ADD(Nz, ch[i].N[2]);
        apx = WN_ADD(apx, ch[i].areaP[0]);
        apy = WN_ADD(apy, ch[i].areaP[1]);
        apz = WN_ADD(apz, ch[i].areaP[2]);
        area = WN_ADD(area, ch[i].area);
        lo0 = wn_min(lo0, ch[i].lo[0]);
        lo1 = wn_min(lo1, ch[i].lo[1]);
        lo2 = wn_min(lo2, ch[i].lo[2]);
        hi0 = wn_max(hi0, ch[i].hi[0]);
        hi1 = wn_max(hi1, ch[i].hi[1]);
        hi2 = wn_max(hi2, ch[i].hi[2]);
    }
    out.N[0] = Nx;
    out.N[1] = Ny;
    out.N[2] = Nz;
    out.areaP[0] = apx;
    out.areaP[1] = apy;
    out.areaP[2] = apz;
    out.area = area;
    out.lo[0] = lo0;
    out.lo[1] = lo1;
    out.lo[2] = lo2;
    out.hi[0] = hi0;
    out.hi[1] = hi1;
    out.hi[2] = hi2;
    float Px, Py, Pz;
    if (area > 0.0f) {
        Px = WN_DIV(apx, area);
        Py = WN_DIV(apy, area);
        Pz = WN_DIV(apz, area);
    } else {
        Px = WN_MUL(0.5f, WN_ADD(lo0, hi0));
        Py = WN_MUL(0.5f, WN_ADD(lo1, hi1));
        Pz = WN_MUL(0.5f, WN_ADD(lo2, hi2));
    }
    out.P[0] = Px;
    out.P[1] = Py;
    out.P[2] = Pz;
    float Nd0 = 0, Nd1 = 0, Nd2 = 0, Nxy = 0, Nyx = 0, Nyz = 0, Nzy = 0, Nzx = 0, Nxz = 0;
    float Td0 = 0, Td1 = 0, Td2 = 0, S = 0, Bxy = 0, Bxz = 0, Byz = 0, Byx = 0, Bzx = 0, Bzy = 0;
    for (int i = 0; i < n; ++i) {
        const WnLocal& c = ch[i];
        const float dx = WN_SUB(c.P[0], Px), dy = WN_SUB(c.P[1], Py), dz = WN_SUB(c.P[2], Pz);
        const float nx = c.N[0], ny = c.N[1], nz = c.N[2];
        // first order
        Nd0 = WN_ADD(Nd0, WN_ADD(c.Nd[0], WN_MUL(nx, dx)));
        Nd1 = WN_ADD(Nd1, WN_ADD(c.Nd[1], WN_MUL(ny, dy)));
        Nd2 = WN_ADD(Nd2, WN_ADD(c.Nd[2], WN_MUL(nz, dz)));
        Nxy = WN_ADD(Nxy, WN_ADD(c.Nxy, WN_MUL(nx, dy)));
        Nyx = WN_ADD(Nyx, WN_ADD(c.Nyx, WN_MUL(ny, dx)));
        Nyz = WN_ADD(Nyz, WN_ADD(c.Nyz, WN_MUL(ny, dz)));
        Nzy = WN_ADD(Nzy, WN_ADD(c.Nzy, WN_MUL(nz, dy)));
        Nzx = WN_ADD(Nzx, WN_ADD(c.Nzx, WN_MUL(nz, dx)));
        Nxz = WN_ADD(Nxz, WN_ADD(c.Nxz, WN_MUL(nx, dz)));
        // second order diagonal: N_iii + 2 d_i N_ii + d_i^2 N_i
        Td0 = WN_ADD(Td0, WN_ADD(WN_ADD(c.Td[0], WN_MUL(2.0f, WN_MUL(dx, c.Nd[0]))), WN_MUL(WN_MUL(dx, dx), nx)));
        Td1 = WN_ADD(Td1, WN_ADD(WN_ADD(c.Td[1], WN_MUL(2.0f, WN_MUL(dy, c.Nd[1]))), WN_MUL(WN_MUL(dy, dy), ny)));
        Td2 = WN_ADD(Td2, WN_ADD(WN_ADD(c.Td[2], WN_MUL(2.0f, WN_MUL(dz, c.Nd[2]))), WN_MUL(WN_MUL(dz, dz), nz)));
        // S + 2 [ dz (Nxy+Nyx) + dy (Nxz+Nzx) + dx (Nyz+Nzy) + nx dy dz + ny dz dx + nz dx dy ]
        float t = WN_MUL(dz, WN_ADD(c.Nxy, c.Nyx));
        t = WN_ADD(t, WN_MUL(dy, WN_ADD(c.Nxz, c.Nzx)));
        t = WN_ADD(t, WN_MUL(dx, WN_ADD(c.Nyz, c.Nzy)));
        t = WN_ADD(t, WN_MUL(WN_MUL(nx, dy), dz));
        t = WN_ADD(t, WN_MUL(WN_MUL(ny, dz), dx));
        t = WN_ADD(t, WN_MUL(WN_MUL(nz, dx), dy));
        S = WN_ADD(S, WN_ADD(c.S, WN_MUL(2.0f, t)));
        Bxy = WN_ADD(Bxy, wn_shift_B(c.Bxy, c.Nd[0], c.Nxy, c.Nyx, nx, ny, dx, dy));
        Bxz = WN_ADD(Bxz, wn_shift_B(c.Bxz, c.Nd[0], c.Nxz, c.Nzx, nx, nz, dx, dz));
        Byz = WN_ADD(Byz, wn_shift_B(c.Byz, c.Nd[1], c.Nyz, c.Nzy, ny, nz, dy, dz));
        Byx = WN_ADD(Byx, wn_shift_B(c.Byx, c.Nd[1], c.Nyx, c.Nxy, ny, nx, dy, dx));
        Bzx = WN_ADD(Bzx, wn_shift_B(c.Bzx, c.Nd[2], c.Nzx, c.Nxz, nz, nx, dz, dx));
        Bzy = WN_ADD(Bzy, wn_shift_B(c.Bzy, c.Nd[2], c.Nzy, c.Nyz, nz, ny, dz, dy));
    }
    out.Nd[0] = Nd0;
    out.Nd[1] = Nd1;
    out.Nd[2] = Nd2;
    out.Nxy = Nxy;
    out.Nyx = Nyx;
    out.Nyz = Nyz;
    out.Nzy = Nzy;
    out.Nzx = Nzx;
    out.Nxz = Nxz;
    out.Td[0] = Td0;
    out.Td[1] = Td1;
    out.Td[2] = Td2;
    out.S = S;
    out.Bxy = Bxy;
    out.Bxz = Bxz;
    out.Byz = Byz;
    out.Byx = Byx;
    out.Bzx = Bzx;
    out.Bzy = Bzy;
    out.pad = 0.0f;
}

// Reference radius (A.3): squared distance from P to the farthest corner of the node's AABB.
WN_HD float wn_box_corner_r2(const WnLocal& d)
{
    const float mx = wn_max(WN_SUB(d.P[0], d.lo[0]), WN_SUB(d.hi[0], d.P[0]));
    const float my = wn_max(WN_SUB(d.P[1], d.lo[1]), WN_SUB(d.hi[1], d.P[1]));
    const float mz = wn_max(WN_SUB(d.P[2], d.lo[2]), WN_SUB(d.hi[2], d.P[2]));
    return WN_ADD(WN_ADD(WN_MUL(mx, mx), WN_MUL(my, my)), WN_MUL(mz, mz));
}

// The reference's stored form (A.4), 23 floats, for parity checks against the oracle's BoxData.
WN_HD void wn_local_to_ref23(const WnLocal& d, float* o)
{
    o[0] = d.P[0];
    o[1] = d.P[1];
    o[2] = d.P[2];
    o[3] = wn_box_corner_r2(d);
    o[4] = d.N[0];
    o[5] = d.N[1];
    o[6] = d.N[2];
    o[7] = d.Nd[0];
    o[8] = d.Nd[1];
    o[9] = d.Nd[2];
    o[10] = WN_ADD(d.Nxy, d.Nyx);
    o[11] = WN_ADD(d.Nyz, d.Nzy);
    o[12] = WN_ADD(d.Nzx, d.Nxz);
    o[13] = d.Td[0];
    o[14] = d.Td[1];
    o[15] = d.Td[2];
    o[16] = d.S;
    o[17] = d.Bxy;
    o[18] = d.Bxz;
    o[19] = d.Byz;
    o[20] = d.Byx;
    o[21] = d.Bzx;
    o[22] = d.Bzy;
}

// Fold the stored form into the evaluation form. With x = r^ (|x| = 1):
//   a0 = -(x . N)
//   a1 = tr(Nij) - 3 x^T Nij x                    = sum_ab q_ab x_a x_b   (trace folded in through |x|^2 = 1)
//   a2 = 1.5 x.(3 D + t0) - 7.5 (cubic terms)      = sum c_abc x_a x_b x_c (linear part folded in through |x|^2 = 1)
//   Omega ~= (a0 + (a1 + a2 / l) / l) / l^2
// order < 2 zeroes the cubic form, order < 1 the quadratic form.
WN_HD void wn_pack_record(const WnLocal& d, float r2, bool leaf, int order, float4* rec /* 6 */)
{
    const float Nxx = d.Nd[0], Nyy = d.Nd[1], Nzz = d.Nd[2];
    const float Cxy = d.Nxy + d.Nyx, Cyz = d.Nyz + d.Nzy, Czx = d.Nzx + d.Nxz;
    const float tr = Nxx + Nyy + Nzz;
    float qxx = tr - 3.0f * Nxx, qyy = tr - 3.0f * Nyy, qzz = tr - 3.0f * Nzz;
    float qxy = -3.0f * Cxy, qyz = -3.0f * Cyz, qzx = -3.0f * Czx;
    const float Lx = 1.5f * (3.0f * d.Td[0] + (d.Byx + d.Bzx));
    const float Ly = 1.5f * (3.0f * d.Td[1] + (d.Bzy + d.Bxy));
    const float Lz = 1.5f * (3.0f * d.Td[2] + (d.Bxz + d.Byz));
    float cxxx = Lx - 7.5f * d.Td[0], cyyy = Ly - 7.5f * d.Td[1], czzz = Lz - 7.5f * d.Td[2];
    float cxyz = -7.5f * d.S;
    float cxxy = Ly - 7.5f * d.Bxy, cxxz = Lz - 7.5f * d.Bxz;
    float cyyz = Lz - 7.5f * d.Byz, cyyx = Lx - 7.5f * d.Byx;
    float czzx = Lx - 7.5f * d.Bzx, czzy = Ly - 7.5f * d.Bzy;
    if (order < 2) cxxx = cyyy = czzz = cxyz = cxxy = cxxz = cyyz = cyyx = czzx = czzy = 0.0f;
    if (order < 1) qxx = qyy = qzz = qxy = qyz = qzx = 0.0f;
    const float r2s = leaf ? wn_int_as_float(wn_float_as_int(r2) | (int)0x80000000) : r2;
    // Storage order = the operand pairs of the evaluation (wn_eval_record): two chains of the same shape advance together,
    // (lo, hi) of one 64-bit register pair, so that sm_100's packed FFMA2/FMUL2 retire two of the record's FMAs per issue slot.
    rec[0] = make_float4(d.P[0], d.P[1], d.P[2], r2s);
    rec[1] = make_float4(czzy, czzz, d.N[0], 0.0f); // .w: link bits, written by the packer
    rec[2] = make_float4(qxx, cxxx, qxy, cxxy);
    rec[3] = make_float4(qzx, cxxz, qyy, cyyx);
    rec[4] = make_float4(qyz, cxyz, qzz, czzx);
    rec[5] = make_float4(d.N[1], cyyy, d.N[2], cyyz);
}

// The 19 expansion coefficients of a stored record in their logical order (tests, debugging):
// N xyz | qxx qyy qzz qxy qyz qzx | cxxx cyyy czzz cxyz cxxy cxxz cyyz cyyx czzx czzy
WN_HD void wn_unpack_record(const float4* rec /* 6 */, float* o /* 19 */)
{
    o[0] = rec[1].z, o[1] = rec[5].x, o[2] = rec[5].z;
    o[3] = rec[2].x, o[4] = rec[3].z, o[5] = rec[4].z, o[6] = rec[2].z, o[7] = rec[4].x, o[8] = rec[3].x;
    o[9] = rec[2].y, o[10] = rec[5].y, o[11] = rec[1].y, o[12] = rec[4].y, o[13] = rec[2].w, o[14] = rec[3].y;
    o[15] = rec[5].w, o[16] = rec[3].w, o[17] = rec[4].w, o[18] = rec[1].x;
}

// ----------------------------------------------------------------------------------------------------------------
// Query arithmetic (fused multiply-adds welcome here)
// ----------------------------------------------------------------------------------------------------------------
WN_HD float wn_rsqrt(float x)
{
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

// One MUFU.RSQ, denormal inputs flushed to zero (-> +inf -> a non-finite expansion -> the caller descends, which is also
// what the reference ends up doing when 1/|r|^2 overflows). Saves the 3-instruction denormal fix-up of rsqrtf().
WN_HD float wn_rsqrt_ftz(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return x < 1.17549435e-38f ? (x == x ? INFINITY : x) : 1.0f / sqrtf(x);
#endif
}

// Far-field Taylor evaluation of one record at r = q - P with l2 = |r|^2 > 0 (A.5 folded, see wn_pack_record).
// f1 = the record's second hot float4, c0..c3 its cold ones. The operation sequence is spelled out (no compiler-chosen
// contraction): this is the definition every kernel variant (scalar here, packed pairs in wn_query.cuh) reproduces bit for bit.
WN_HD float wn_eval_record(float rx, float ry, float rz, float l2, const float4& f1, const float4& c0, const float4& c1,
                           const float4& c2, const float4& c3)
{
    const float m1 = wn_rsqrt_ftz(l2);
    const float x = WN_MUL(rx, m1), y = WN_MUL(ry, m1), z = WN_MUL(rz, m1);
    const float m2 = WN_MUL(m1, m1);
    // (t, p) = x (qxx, cxxx) + y (qxy, cxxy) + z (qzx, cxxz)
    const float t = WN_FMA(z, c1.x, WN_FMA(y, c0.z, WN_MUL(x, c0.x)));
    const float p = WN_FMA(z, c1.y, WN_FMA(y, c0.w, WN_MUL(x, c0.y)));
    // (u, q) = y (qyy, cyyx) + z (qyz, cxyz);  (w, s) = z (qzz, czzx)
    const float u = WN_FMA(z, c2.x, WN_MUL(y, c1.z));
    const float q = WN_FMA(z, c2.y, WN_MUL(y, c1.w));
    const float w = WN_MUL(z, c2.z);
    const float s = WN_MUL(z, c2.w);
    // (a1, cx) = x (t, p) + y (u, q) + z (w, s): the quadratic form and the x-part of the cubic form
    const float a1 = WN_FMA(z, w, WN_FMA(y, u, WN_MUL(x, t)));
    const float cx = WN_FMA(z, s, WN_FMA(y, q, WN_MUL(x, p)));
    // (n, g) = y (Ny, cyyy) + z (Nz, cyyz);  n += x Nx
    const float n = WN_FMA(x, f1.z, WN_FMA(z, c3.z, WN_MUL(y, c3.x)));
    const float g = WN_FMA(z, c3.w, WN_MUL(y, c3.y));
    // (h, k) = z (czzy, czzz)
    const float h = WN_MUL(z, f1.x);
    const float k = WN_MUL(z, f1.y);
    const float cy = WN_FMA(z, h, WN_MUL(y, g));
    const float cz = WN_MUL(z, k);
    const float a2 = WN_FMA(z, cz, WN_FMA(y, cy, WN_MUL(x, cx)));
    // Omega ~= m2 (-x.N + m1 (a1 + m1 a2))
    return WN_MUL(m2, WN_FMA(m1, WN_FMA(m1, a2, a1), -n));
}

// Exact signed solid angle of triangle (a,b,c) seen from q (A.1, reference formulation incl. its two zero rules).
WN_HD float wn_tri_solid_angle(float qx, float qy, float qz, const float4& a, const float4& b, const float4& c)
{
    float ax = a.x - qx, ay = a.y - qy, az = a.z - qz;
    float bx = b.x - qx, by = b.y - qy, bz = b.z - qz;
    float cx = c.x - qx, cy = c.y - qy, cz = c.z - qz;
    const float al2 = ax * ax + ay * ay + az * az;
    const float bl2 = bx * bx + by * by + bz * bz;
    const float cl2 = cx * cx + cy * cy + cz * cz;
    if (al2 == 0.0f || bl2 == 0.0f || cl2 == 0.0f) return 0.0f;
    const float ia = wn_rsqrt(al2), ib = wn_rsqrt(bl2), ic = wn_rsqrt(cl2);
    ax *= ia;
    ay *= ia;
    az *= ia;
    bx *= ib;
    by *= ib;
    bz *= ib;
    cx *= ic;
    cy *= ic;
    cz *= ic;
    const float ux = bx - ax, uy = by - ay, uz = bz - az;
    const float vx = cx - ax, vy = cy - ay, vz = cz - az;
    const float num = ax * (uy * vz - uz * vy) + ay * (uz * vx - ux * vz) + az * (ux * vy - uy * vx);
    if (num == 0.0f) return 0.0f;
    const float den = 1.0f + (ax * bx + ay * by + az * bz) + (ax * cx + ay * cy + az * cz) + (bx * cx + by * cy + bz * cz);
    return 2.0f * atan2f(num, den);
}

// Exact mode, fast path. The half solid angle of a triangle is the argument of the complex number (den, num) of
// wn_tri_solid_angle, and arguments add under multiplication: a group of triangles that each subtend a small angle contributes
// arg(prod (den_k + i num_k)) — one atan2 per group instead of one per triangle. This folds one triangle into the running product
// z and reports whether it qualifies: |half angle| < pi/8 (so that eight of them cannot wrap) and den >= 1/2 (so that the product
// of eight cannot underflow). Anything else (a triangle the query is close to, a vertex on the query, NaN) makes the caller
// redo the group term by term with wn_tri_solid_angle, which carries the reference's zero rules. Same normalised formulation
// (one MUFU.RSQ each, denormal squared lengths flush to +inf -> NaN -> the careful path).
#define WN_EXACT_GROUP 8
WN_HD bool wn_tri_fold(float qx, float qy, float qz, const float4& a, const float4& b, const float4& c, float& zr, float& zi)
{
    float ax = a.x - qx, ay = a.y - qy, az = a.z - qz;
    float bx = b.x - qx, by = b.y - qy, bz = b.z - qz;
    float cx = c.x - qx, cy = c.y - qy, cz = c.z - qz;
    const float ia = wn_rsqrt_ftz(ax * ax + ay * ay + az * az), ib = wn_rsqrt_ftz(bx * bx + by * by + bz * bz), ic = wn_rsqrt_ftz(cx * cx + cy * cy + cz * cz);
    ax *= ia;
    ay *= ia;
    az *= ia;
    bx *= ib;
    by *= ib;
    bz *= ib;
    cx *= ic;
    cy *= ic;
    cz *= ic;
    const float ux = bx - ax, uy = by - ay, uz = bz - az;
    const float vx = cx - ax, vy = cy - ay, vz = cz - az;
    const float num = ax * (uy * vz - uz * vy) + ay * (uz * vx - ux * vz) + az * (ux * vy - uy * vx);
    const float den = 1.0f + (ax * bx + ay * by + az * bz) + (ax * cx + ay * cy + az * cz) + (bx * cx + by * cy + bz * cz);
    const float nr = zr * den - zi * num, ni = zr * num + zi * den;
    zr = nr;
    zi = ni;
    // tan(pi/8) = 0.41421356; a hair less, so that eight half angles stay strictly inside (-pi, pi). False for NaN.
    return fabsf(num) <= 0.4142f * den && den >= 0.5f;
}

// Closest point of triangle (a, b, c) to p and its squared distance: closest-feature regions (Ericson, Real-Time Collision
// Detection, 5.1.5). What TriangleAABBTree::get_closest_point evaluates per candidate triangle
// (modules/bvh/include/lagrange/bvh/TriangleAABBTree.h:84-88).
WN_HD float wn_point_tri_closest(float px, float py, float pz, const float4& a, const float4& b, const float4& c, float& ox, float& oy, float& oz)
{
    const float abx = b.x - a.x, aby = b.y - a.y, abz = b.z - a.z;
    const float acx = c.x - a.x, acy = c.y - a.y, acz = c.z - a.z;
    const float apx = px - a.x, apy = py - a.y, apz = pz - a.z;
    const float d1 = abx * apx + aby * apy + abz * apz;
    const float d2 = acx * apx + acy * apy + acz * apz;
    float cx, cy, cz; // closest point
    if (d1 <= 0.0f && d2 <= 0.0f) {
        cx = a.x, cy = a.y, cz = a.z;
    } else {
        const float bpx = px - b.x, bpy = py - b.y, bpz = pz - b.z;
        const float d3 = abx * bpx + aby * bpy + abz * bpz;
        const float d4 = acx * bpx + acy * bpy + acz * bpz;
        if (d3 >= 0.0f && d4 <= d3) {
            cx = b.x, cy = b.y, cz = b.z;
        } else {
            const float vc = d1 * d4 - d3 * d2;
            if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
                const float v = d1 / (d1 - d3);
                cx = a.x + v * abx, cy = a.y + v * aby, cz = a.z + v * abz;
            } else {
                const float cpx = px - c.x, cpy = py - c.y, cpz = pz - c.z;
                const float d5 = abx * cpx + aby * cpy + abz * cpz;
                const float d6 = acx * cpx + acy * cpy + acz * cpz;
                if (d6 >= 0.0f && d5 <= d6) {
                    cx = c.x, cy = c.y, cz = c.z;
                } else {
                    const float vb = d5 * d2 - d1 * d6;
                    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
                        const float w = d2 / (d2 - d6);
                        cx = a.x + w * acx, cy = a.y + w * acy, cz = a.z + w * acz;
                    } else {
                        const float va = d3 * d6 - d5 * d4;
                        if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
                            const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
                            cx = b.x + w * (c.x - b.x), cy = b.y + w * (c.y - b.y), cz = b.z + w * (c.z - b.z);
                        } else {
                            const float denom = 1.0f / (va + vb + vc);
                            const float v = vb * denom, w = vc * denom;
                            cx = a.x + abx * v + acx * w, cy = a.y + aby * v + acy * w, cz = a.z + abz * v + acz * w;
                        }
                    }
                }
            }
        }
    }
    const float dx = px - cx, dy = py - cy, dz = pz - cz;
    ox = cx, oy = cy, oz = cz;
    return dx * dx + dy * dy + dz * dz;
}
WN_HD float wn_point_tri_dist2(float px, float py, float pz, const float4& a, const float4& b, const float4& c)
{
    float x, y, z;
    return wn_point_tri_closest(px, py, pz, a, b, c, x, y, z);
}

// FastWindingNumber.cpp:66 computes  omega / (4.f * pi) > 0.5f  with pi a double (constants.h:16), i.e. in double.
// Over all floats this is exactly  omega >= 6.2831854820251465f  (the float nearest to, and just above, 2 pi);
// tests/test_predicate.py checks the equivalence against the oracle's literal restatement.
#define WN_INSIDE_THRESHOLD 6.2831854820251465f
WN_HD bool wn_inside_from_omega(float omega)
{
    return omega >= WN_INSIDE_THRESHOLD;
}

// Cell-centred lattice point (mesh_to_volume.cpp:147-149), float arithmetic, no contraction so host and device agree.
WN_HD float wn_lattice_coord(float origin, float spacing, int i)
{
    return WN_ADD(origin, WN_MUL(spacing, WN_ADD((float)i, 0.5f)));
}

// One query point against the packed tree, one lane, no warp cooperation: the definition of the per-point result.
// (Used by the host emulation harness and mirrored lane-wise by the warp kernel.)
// Memory layout of the packed tree: the six float4 of a record (see wn_pack_record) are split in a hot and a cold part,
// each interleaved per entry, so that one address computation serves every load of a visit:
//   hot [2*i + 0] = (Px, Py, Pz, R2 | leaf sign)     hot [2*i + 1] = (czzy, czzz, Nx, link bits)   32 B: one sector
//   cold[4*i + k] = the paired form coefficients (rec[2..5])                                        64 B: two sectors
// A visit that only tests (the record is near for every lane) touches the hot sector alone.
struct WnTreeView
{
    const float4* hot;
    const float4* cold;
    const int4* kids;  // entry indices of an internal entry's children (-1 = none); lets the tile planner expand a node
                       // without walking the sibling chain through dependent link loads
    const float4* tri; // 3 float4 per triangle: a, b, c (w unused), depth-first order
    int n_entries;
    int n_tris;
};

WN_HD float wn_traverse_point(const WnTreeView& t, float qx, float qy, float qz, float beta2, unsigned long long* cnt /* 3 or null */)
{
    float acc = 0.0f;
    // entry 0 is the root, which is never approximated (A.5): start at its first child unless the root is itself a leaf
    int i = t.n_entries > 1 ? 1 : 0;
    while (i < t.n_entries) {
        const float4 f0 = t.hot[2 * (int64_t)i], f1 = t.hot[2 * (int64_t)i + 1];
        const int lk = wn_float_as_int(f1.w);
        const bool leaf = wn_float_as_int(f0.w) < 0;
        const float thr = WN_MUL(fabsf(f0.w), beta2);
        const float rx = qx - f0.x, ry = qy - f0.y, rz = qz - f0.z;
        const float l2 = WN_ADD(WN_ADD(WN_MUL(rx, rx), WN_MUL(ry, ry)), WN_MUL(rz, rz)); // unfused, like the reference
        bool near = l2 <= thr;
        if (cnt) cnt[0]++;
        if (!near) {
            const float4* c = t.cold + 4 * (int64_t)i;
            const float om = wn_eval_record(rx, ry, rz, l2, f1, c[0], c[1], c[2], c[3]);
            if (fabsf(om) <= 3.402823466e38f) { // finite
                acc += om;
                if (cnt) cnt[1]++;
                i = leaf ? i + 1 : lk;
                continue;
            }
            near = true;
        }
        if (leaf) {
            const int first = lk >> WN_LEAF_COUNT_BITS, count = (lk & (WN_MAX_LEAF_SIZE - 1)) + 1;
            for (int k = 0; k < count; ++k) {
                acc += wn_tri_solid_angle(qx, qy, qz, t.tri[3 * (first + k)], t.tri[3 * (first + k) + 1], t.tri[3 * (first + k) + 2]);
                if (cnt) cnt[2]++;
            }
        }
        i = i + 1;
    }
    return acc;
}
