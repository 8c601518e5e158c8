// wn_packed.h — layout of the packed tree (one position-independent allocation, header first). Shared by the CUDA library
// (wn_capi.cu: build, wn_tree_pack, wn_create_from_packed) and the C++ host layer (src/FastWindingNumber.cpp: the
// reference-surface single-point overloads walk a HOST copy of this blob).
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <string.h>

struct WnPackedHeader
{
    uint64_t magic; // WN_PACKED_MAGIC
    int64_t total_bytes;
    int64_t n_entries;
    int64_t n_tris;
    int64_t off_hot;  // float4[2 * n_entries]: (P, R2 | leaf), (czzy, czzz, Nx, link bits)
    int64_t off_cold; // float4[4 * n_entries]: quadratic + cubic form in operand-pair order (wn_pack_record)
    int64_t off_kids; // int4[n_entries]: entry indices of an internal entry's children
    int64_t off_tris; // float4[3 * n_tris]: triangles in depth-first order
    int64_t off_tri_order; // uint32[n_tris]: input triangle id at each depth-first position
    int32_t width;
    int32_t order;
    float accuracy_scale;
    int32_t max_depth;
    int64_t n_leaf_entries;
    int64_t num_vertices;
    int64_t num_tree_nodes;
    int64_t reserved[3];
};
#define WN_PACKED_MAGIC 0x3454303032424e57ull /* 'WNB200T4' (T4: paired coefficient order, wn_pack_record) */

static inline size_t wn_packed_align(size_t v, size_t a)
{
    return (v + a - 1) / a * a;
}

static inline WnPackedHeader wn_packed_make_header(int64_t n_entries, int64_t n_tris)
{
    WnPackedHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = WN_PACKED_MAGIC;
    size_t off = wn_packed_align(sizeof(WnPackedHeader), 256);
    h.off_hot = (int64_t)off;
    off = wn_packed_align(off + (size_t)n_entries * 2 * 16, 256);
    h.off_cold = (int64_t)off;
    off = wn_packed_align(off + (size_t)n_entries * 4 * 16, 256);
    h.off_kids = (int64_t)off;
    off = wn_packed_align(off + (size_t)n_entries * 16, 256);
    h.off_tris = (int64_t)off;
    off = wn_packed_align(off + (size_t)n_tris * 3 * 16, 256);
    h.off_tri_order = (int64_t)off;
    off = wn_packed_align(off + (size_t)n_tris * sizeof(uint32_t), 256);
    h.total_bytes = (int64_t)off;
    h.n_entries = n_entries;
    h.n_tris = n_tris;
    return h;
}
