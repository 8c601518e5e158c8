// wn_sdf.cuh — K10: narrow-band distance to the mesh on a cell-centred lattice, signed by the winding number.
// The GPU side of the reference's one production caller, volume::mesh_to_volume with Sign::WindingNumber
// (modules/volume/src/mesh_to_volume.cpp:147-183: OpenVDB meshToVolume, exterior = interior band width = 3 voxels, the
// interior test is FastWindingNumber::is_inside at every voxel centre). SURVEY.md section 8(f) N1/N3.
//
// Distance: closest-triangle search on the same packed hierarchy the winding-number queries walk. Every record already
// carries a bounding sphere of its subtree (expansion centre P, box-corner radius R: every triangle of the node lies within
// R of P), so a subtree is culled when |q - P| - R exceeds the best distance so far, which starts at the band width: only
// the hierarchy near the voxel is ever opened, and voxels outside the band stop after a few tests. Same warp-cooperative
// stackless walk as warp_traverse: the warp opens a subtree when any lane needs it.
#pragma once

#include "wn_query.cuh"

namespace wn {

struct SdfArgs
{
    WnTreeView tree;
    GridDesc g;
    int tiles_x, tiles_y;
    float band;            // world units; distances are clamped to it
    const uint8_t* inside; // [nz*ny*nx] winding-number sign per voxel, or null: unsigned distance
    float* out;            // [nz*ny*nx] signed distance (negative inside), +-band outside the narrow band
    unsigned long long* active; // optional: number of voxels with |d| < band
};

// CTA = 8x8x8 voxels, warp = 4x4x4, two voxels per lane (z and z+2), like k_query<2, GRID>.
__global__ void __launch_bounds__(kQueryThreads) k_sdf_grid(const SdfArgs a)
{
    constexpr int QPL = 2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bx = blockIdx.x % a.tiles_x, by = (blockIdx.x / a.tiles_x) % a.tiles_y, bz = blockIdx.x / (a.tiles_x * a.tiles_y);
    const int x = bx * 8 + (wid & 1) * 4 + (lane & 3);
    const int y = by * 8 + ((wid >> 1) & 1) * 4 + ((lane >> 2) & 3);
    float qx[QPL], qy[QPL], qz[QPL], best[QPL], best2[QPL];
    bool valid[QPL];
    int64_t oidx[QPL];
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        const int z = bz * 8 + (wid >> 2) * 4 + 2 * k + (lane >> 4);
        valid[k] = x < a.g.nx && y < a.g.ny && z < a.g.nz;
        qx[k] = wn_lattice_coord(a.g.ox, a.g.sx, x);
        qy[k] = wn_lattice_coord(a.g.oy, a.g.sy, y);
        qz[k] = wn_lattice_coord(a.g.oz, a.g.sz, z);
        oidx[k] = valid[k] ? ((int64_t)z * a.g.ny + y) * a.g.nx + x : -1;
        best[k] = a.band;
        best2[k] = a.band * a.band;
    }
    const float4* __restrict__ hot = a.tree.hot;
    const float4* __restrict__ tris = a.tree.tri;
    const int n = a.tree.n_entries;
    int i = 0;
    while (i < n) {
        const float4 f0 = __ldg(hot + 2 * (int64_t)i);
        const int lk = __float_as_int(__ldg(&hot[2 * (int64_t)i + 1].w));
        const bool leaf = __float_as_int(f0.w) < 0;
        const float R = sqrtf(fabsf(f0.w)); // +inf for records that are never approximated: never culled
        bool need[QPL], any = false;
#pragma unroll
        for (int k = 0; k < QPL; ++k) {
            const float rx = qx[k] - f0.x, ry = qy[k] - f0.y, rz = qz[k] - f0.z;
            const float d2 = rx * rx + ry * ry + rz * rz;
            const float reach = R + best[k];
            // conservative: a subtree is dropped only if it is clearly out of reach (1e-5 relative slack for rounding)
            need[k] = valid[k] && !(d2 > reach * reach * 1.00001f);
            any |= need[k];
        }
        if (!__any_sync(kFull, any)) {
            i = leaf ? i + 1 : lk;
            continue;
        }
        if (leaf) {
            const int first = lk >> WN_LEAF_COUNT_BITS, count = (lk & (WN_MAX_LEAF_SIZE - 1)) + 1;
            for (int tt = 0; tt < count; ++tt) {
                const float4 ta = __ldg(tris + 3 * (int64_t)(first + tt));
                const float4 tb = __ldg(tris + 3 * (int64_t)(first + tt) + 1);
                const float4 tc = __ldg(tris + 3 * (int64_t)(first + tt) + 2);
#pragma unroll
                for (int k = 0; k < QPL; ++k) {
                    if (need[k]) {
                        const float d2 = wn_point_tri_dist2(qx[k], qy[k], qz[k], ta, tb, tc);
                        if (d2 < best2[k]) {
                            best2[k] = d2;
                            best[k] = sqrtf(d2);
                        }
                    }
                }
            }
        }
        i = i + 1;
    }
    unsigned int act = 0;
#pragma unroll
    for (int k = 0; k < QPL; ++k) {
        if (oidx[k] >= 0) {
            const bool in = a.inside ? a.inside[oidx[k]] != 0 : false;
            a.out[oidx[k]] = in ? -best[k] : best[k];
            act += best[k] < a.band ? 1u : 0u;
        }
    }
    if (a.active) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) act += __shfl_xor_sync(kFull, act, o);
        if (lane == 0 && act) atomicAdd(a.active, (unsigned long long)act);
    }
}

} // namespace wn
