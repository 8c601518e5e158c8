/*
 * wn_b200.h — C-ABI of the B200-native fast generalized winding number engine (libwn_b200.so).
 *
 * This is the drop-in boundary for the lagrange::winding::FastWindingNumber hot path: plain pointers and sizes, no
 * C++/torch types. Every entry point names the reference interface it replaces (paths relative to the adobe/lagrange
 * checkout). The engine behind lagrange's wrapper is HDK_Sample::UT_SolidAngle<float,float> (third-party, fetched by
 * cmake/recipes/external/winding_number.cmake:21-26); lagrange binds exactly two of its methods:
 *
 *     m_engine.init(num_triangles, triangles_ptr, num_vertices, m_vertices.data());   FastWindingNumber.cpp:57
 *     m_engine.computeSolidAngle(q)  [accuracy_scale = 2, order = 2 defaults]         FastWindingNumber.cpp:66,75
 *
 * INTEGRATION.md shows the binding a lagrange maintainer would write over these functions.
 *
 * Conventions
 *   - All functions return wn_status (0 = ok). On failure wn_last_error() returns a thread-local message.
 *   - Pointers named *_xyz hold packed float triples. Data pointers may be host or device pointers; residency is
 *     detected with cudaPointerGetAttributes. Host buffers are staged through pinned memory inside the call.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream). Calls with device buffers are
 *     asynchronous with respect to the host on that stream; calls with host buffers return after the results landed.
 *   - Query entry points are re-entrant: one engine may be queried from many host threads (the reference is called
 *     from OpenVDB's TBB workers, modules/volume/src/mesh_to_volume.cpp:175-183); calls serialise per engine.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with WN_ERR_CUDA.
 */
#ifndef WN_B200_H
#define WN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define WN_API __declspec(dllexport)
#else
#define WN_API __attribute__((visibility("default")))
#endif

typedef struct wn_engine wn_engine;

typedef enum wn_status {
    WN_OK = 0,
    WN_ERR_INVALID_ARGUMENT = 1, /* null pointer, negative size, index out of range, malformed topology */
    WN_ERR_CUDA = 2,             /* CUDA runtime/driver error, or no device */
    WN_ERR_OUT_OF_MEMORY = 3,
    WN_ERR_UNSUPPORTED = 4       /* e.g. more than 2^27 triangles */
} wn_status;

/* Radius used by the far-field acceptance test  |q - P|^2 > beta^2 * R^2. */
typedef enum wn_radius_mode {
    WN_RADIUS_BOX_CORNER = 0, /* reference: distance from the area-weighted centroid to the farthest AABB corner */
    WN_RADIUS_VERTEX = 1      /* exact: distance to the farthest vertex of the cluster (never larger than BOX_CORNER) */
} wn_radius_mode;

/* Hierarchy built by wn_create (both on the GPU; moments, packing and traversal are shared). */
typedef enum wn_hierarchy {
    WN_HIERARCHY_LBVH = 0, /* Morton codes + one radix sort + Karras 2012: fastest build (1.9 ms for 1.3 M triangles) */
    WN_HIERARCHY_KD = 1,   /* balanced k-d: object-median splits along the longest centroid axis, one radix sort per
                              level: compact equal-count patches, ~13 % faster queries, ~4 ms more build time */
    WN_HIERARCHY_KD_SAH = 2, /* k-d with the split position chosen by the surface-area heuristic among the boundaries of 16
                              equal-width bins along the axis (the reference builder's rule): fastest queries, ~11 ms build */
    WN_HIERARCHY_REFERENCE = 3 /* the reference builder's own tree (UT_BVH<4>::init<BOX_AREA>: 4-ary, top-down, binned SAH with its
                              small-range rules, one triangle per leaf slot), reproduced on the GPU bit for bit against the CPU
                              restatement: solid_angle then matches the reference algorithm to float rounding (<= 1e-4 * 4 pi)
                              and is_inside is identical outside |w - 0.5| <= 1e-3. leaf_size is forced to 1 and single
                              triangles are approximated by their own expansion like the reference does */
} wn_hierarchy;

typedef struct wn_options {
    uint32_t struct_size;     /* = sizeof(wn_options); set by wn_options_init */
    int32_t device;           /* CUDA device ordinal, -1 = current device */
    float accuracy_scale;     /* default beta for queries that pass beta <= 0; reference default 2 (UT_SolidAngle) */
    int32_t order;            /* Taylor order of the far field: 0, 1 or 2; reference default 2 */
    int32_t leaf_size;        /* LBVH build: max triangles per leaf (1..16). Imported topologies are always 1 */
    int32_t morton_bits;      /* LBVH build: 30 or 63 */
    int32_t radius_mode;      /* wn_radius_mode */
    int32_t approximate_single_triangles; /* 1 = far single-triangle leaves use their dipole expansion like the
                                             reference (A.4: "single-triangle children get a record too"); 0 = they are
                                             always evaluated exactly (cheaper and more accurate). Imported topologies
                                             default to 1, the LBVH build to 0 */
    int32_t keep_build_data;  /* 1 = keep per-node raw moments for wn_debug_node_moments */
    int32_t hierarchy;        /* wn_hierarchy (default WN_HIERARCHY_REFERENCE); ignored by wn_create_from_topology */
    int32_t reserved[6];
} wn_options;

typedef struct wn_info {
    uint32_t struct_size;
    int32_t device;
    int64_t num_vertices;
    int64_t num_triangles;
    int64_t num_tree_nodes;      /* nodes of the hierarchy (internal + leaves) before packing */
    int64_t num_entries;         /* records in the packed depth-first array the traversal walks (incl. root) */
    int64_t num_leaf_entries;
    int64_t tree_bytes;          /* bytes of the packed tree (node records + links + triangle records) */
    int64_t build_scratch_bytes; /* peak temporary device memory used by the build */
    float build_ms;              /* device time of the whole build (CUDA events) */
    float build_ms_morton;       /* K1 bounds + Morton codes */
    float build_ms_sort;         /* K2 radix sort */
    float build_ms_hierarchy;    /* K3 LBVH topology (0 for imported topologies) */
    float build_ms_moments;      /* K4 bottom-up moments */
    float build_ms_pack;         /* K5 depth-first packing */
    int32_t max_depth;
    int32_t width;               /* 2 = LBVH, 4 = imported UT_BVH<4>-style topology */
    float accuracy_scale;
    int32_t order;
} wn_info;

/* Work EXECUTED by one batch of tree queries, in the units of SURVEY.md section 8(d):
 * flops = 10 * node_tests + 83 * far_field_evals + 75 * exact_triangles.
 * With WN_QUERY_NO_TILING these are the per-point counts of the reference algorithm (every point tests / expands /
 * descends exactly like UT_SolidAngle::computeSolidAngle does on the same tree). The tiled path accepts the same
 * records per point but executes less: records that are far for a whole tile are evaluated at 64 sample points per
 * tile (counted in far_field_evals) and records that are near for a whole tile are not tested at all. */
typedef struct wn_query_stats {
    uint64_t node_tests;
    uint64_t far_field_evals;
    uint64_t exact_triangles;
    uint64_t lane_slots;         /* (lane, query) slots the warps stepped through = 32 * QPL per visited record;
                                    lane utilisation of the traversal = node_tests / lane_slots */
} wn_query_stats;

WN_API const char* wn_last_error(void);
WN_API const char* wn_version(void);
WN_API wn_status wn_options_init(wn_options* opt);

/* ---- construction -------------------------------------------------------------------------------------------
 * Replaces UT_SolidAngle::init(ntris, tri_pts, npts, positions) as called at FastWindingNumber.cpp:54-57.
 * Input buffers are copied (the reference keeps converted copies alive for the same reason, :40-52,82-84); the
 * caller may free them when the call returns. GPU build: K1 bounds + Morton codes, K2 radix sort, K3 LBVH, K4
 * bottom-up order-2 moments, K5 depth-first packing. nT == 0 builds an empty engine (Omega = 0 everywhere). */
WN_API wn_status wn_create(const float* v_xyz, int64_t num_vertices, const int32_t* tri, int64_t num_triangles,
                           const wn_options* opt, wn_engine** out);

/* Same, but the hierarchy topology is supplied by the caller ("oracle-tree mode"): `child` holds
 * num_nodes * width slots (width 2..4), node 0 is the root, slot encoding: c >= 0 internal node index, c == -1 empty
 * (trailing), c <= -2 triangle index -(c+2). Every triangle must appear exactly once. Moments, radii and packing
 * are still computed on the GPU (K4, K5). With a UT_BVH<4> topology this reproduces the reference's tree, so query
 * results agree with the reference algorithm to float rounding. */
WN_API wn_status wn_create_from_topology(const float* v_xyz, int64_t num_vertices, const int32_t* tri, int64_t num_triangles,
                                         const int32_t* child, int64_t num_nodes, int32_t width, const wn_options* opt,
                                         wn_engine** out);

WN_API wn_status wn_destroy(wn_engine* e);
WN_API wn_status wn_get_info(const wn_engine* e, wn_info* info);

/* ---- tree queries (K6) --------------------------------------------------------------------------------------
 * Replace UT_SolidAngle::computeSolidAngle(q, accuracy_scale) (FastWindingNumber.cpp:66,75), batched.
 * beta <= 0 selects the engine default. `flags`: see WN_QUERY_*. */
#define WN_QUERY_DEFAULT 0u
#define WN_QUERY_PRESORTED 1u /* points are already spatially coherent: skip the Morton sort of the queries (K9) */
#define WN_QUERY_NO_TILING 2u /* always run the generic per-point traversal (no tile plan, no far-field interpolation) */
#define WN_QUERY_OUT_BITS 8u  /* out_inside is a bit array: query i -> bit (i & 7) of byte i >> 3, (n + 7) / 8 bytes (numpy:
                                 unpackbits(bitorder='little')). 8x less device-to-host traffic for host outputs. wn_is_inside,
                                 wn_query_grid, wn_query_grid_strided */

/* out_omega[i] = solid angle at q_i, in (-4pi k, 4pi k): what FastWindingNumber::solid_angle returns (:69-76). */
WN_API wn_status wn_solid_angle(const wn_engine* e, const float* q_xyz, int64_t n, float beta, uint32_t flags,
                                float* out_omega, void* stream);
/* out_inside[i] = 1 iff (double)omega / (4.0 * pi) > 0.5 : FastWindingNumber::is_inside (:60-67). */
WN_API wn_status wn_is_inside(const wn_engine* e, const float* q_xyz, int64_t n, float beta, uint32_t flags,
                              uint8_t* out_inside, void* stream);

/* Implicit cell-centred lattice, the mesh_to_volume call pattern (modules/volume/src/mesh_to_volume.cpp:147-149,
 * 175-182): p(i,j,k) = origin + spacing * ((i,j,k) + 0.5), evaluated in float; x fastest, then y, then z.
 * Only the z-slab [z_begin, z_end) is evaluated (multi-GPU sharding); the output holds dims[0]*dims[1]*(z_end-z_begin)
 * values starting at the slab's first point. Either output may be NULL (not both). */
WN_API wn_status wn_query_grid(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3],
                               int64_t z_begin, int64_t z_end, float beta, uint32_t flags, float* out_omega, uint8_t* out_inside,
                               void* stream);

/* Strided sharding of a lattice across GPUs: evaluates the tile layers (8 consecutive z-planes each, counted from z = 0)
 * layer_first, layer_first + layer_step, layer_first + 2*layer_step, ... and stores them compactly in that order (each
 * layer dims[0]*dims[1]*8 values, the lattice's last layer possibly fewer). With layer_step = number of ranks and
 * layer_first = rank every rank gets the same mix of work in one call (contiguous z-slabs load-imbalance when the work is
 * not uniform along z). */
WN_API wn_status wn_query_grid_strided(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3],
                                       int64_t layer_first, int64_t layer_step, float beta, uint32_t flags, float* out_omega,
                                       uint8_t* out_inside, void* stream);

/* Diagonal sharding of a lattice across `world` GPUs: the lattice is cut in Q parts along y (Q = 4, 2 or 1: the
 * largest that divides both `world` and the tile rows, ny % (8 Q) == 0) and rank r evaluates, of every c-th tile layer (c = world / Q,
 * starting at layer r mod c), the one part ((r - layer) / c) mod Q. Every rank sees every part and every height equally often and the
 * unit of balance is a Q-th of a layer: for lattices with few tile layers per rank. (bench.py deals whole layers,
 * wn_query_grid_strided: on the 64-layer cfg2 lattice both measured the same within 1 %.) Output: the rank's units in layer order, each planes x part_rows x nx values (planes = 8, fewer for the last
 * layer); wn_grid_shard_layout describes it. Q == 1 is exactly wn_query_grid_strided(rank, world). Honours WN_QUERY_OUT_BITS. */
WN_API wn_status wn_query_grid_sharded(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3], int32_t rank,
                                       int32_t world, float beta, uint32_t flags, float* out_omega, uint8_t* out_inside, void* stream);
/* Layout of rank's output: unit k (k = 0 .. n_units-1) = tile layer lz = rank % layer_step + k * layer_step, rows
 * [q * part_rows, (q + 1) * part_rows) with q = ((rank - lz) / layer_step) mod parts_y (non-negative), planes [8 lz, min(nz, 8 lz + 8)). */
WN_API wn_status wn_grid_shard_layout(const int64_t dims[3], int32_t rank, int32_t world, int32_t* parts_y, int32_t* layer_step,
                                      int64_t* part_rows, int64_t* n_units, int64_t* n_points);

/* Counters of the traversal for a batch of points (same traversal as wn_solid_angle, results discarded). */
WN_API wn_status wn_query_stats_points(const wn_engine* e, const float* q_xyz, int64_t n, float beta, uint32_t flags,
                                       wn_query_stats* stats, void* stream);
WN_API wn_status wn_query_stats_grid(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3],
                                     int64_t z_begin, int64_t z_end, float beta, uint32_t flags, wn_query_stats* stats, void* stream);

/* ---- exact brute-force mode (K7) ----------------------------------------------------------------------------
 * Sum of exact Van Oosterom-Strackee triangle solid angles over ALL triangles (UTsignedSolidAngleTri):
 * tree-independent ground truth on the device. Either output may be NULL (not both). */
WN_API wn_status wn_exact(const wn_engine* e, const float* q_xyz, int64_t n, float* out_omega, uint8_t* out_inside, void* stream);
WN_API wn_status wn_exact_grid(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3],
                               int64_t z_begin, int64_t z_end, float* out_omega, uint8_t* out_inside, void* stream);

/* ---- narrow-band signed distance on a lattice (replaces the OpenVDB meshToVolume call of -------
 * volume::mesh_to_volume with Sign::WindingNumber, modules/volume/src/mesh_to_volume.cpp:147-183) ----------------------
 * out_sdf[(z*ny + y)*nx + x] = s * min(distance from the cell centre origin + spacing*(ijk + 1/2) to the mesh, band), with
 * s = -1 where the winding-number predicate (wn_is_inside, accuracy `beta`) holds and +1 elsewhere; WN_SDF_UNSIGNED skips
 * the predicate (s = +1). band is in world units (the reference uses 3 voxels on either side). The distance is exact:
 * closest triangle, searched on the engine's own hierarchy. out_sdf: host or device. num_active (optional, host): number of
 * cells with distance < band (the narrow band OpenVDB would keep active). */
#define WN_SDF_UNSIGNED 4u
WN_API wn_status wn_sdf_grid(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3], float band,
                             float beta, uint32_t flags, float* out_sdf, int64_t* num_active, void* stream);

/* Same, sparse: only the cells of the narrow band (|d| < band) are returned, ordered by linear index (z*ny + y)*nx + x — what an
 * OpenVDB FloatGrid keeps active (mesh_to_volume.cpp:160-183); the dense block stays in device scratch. Up to `capacity` cells
 * are written to out_index / out_value (host or device); *num_active is the size of the band (call with capacity = 0 to size
 * the buffers). out_inside_bits (optional, (n + 7) / 8 bytes): the interior test of EVERY cell, 1 bit each, for the sign of the
 * inactive interior. 512^3 voxels around a 1.3 M-triangle sphere: ~4 M active cells, 50 MB instead of 537 MB to the host. */
WN_API wn_status wn_sdf_grid_sparse(const wn_engine* e, const float origin[3], const float spacing[3], const int64_t dims[3], float band,
                                    float beta, uint32_t flags, int64_t capacity, int64_t* out_index, float* out_value, uint8_t* out_inside_bits,
                                    int64_t* num_active, void* stream);

/* ---- closest point on the mesh ----------------------------------------------------------------------------
 * Replaces TriangleAABBTree::get_closest_point(p, triangle_id, closest_point, closest_sq_dist)
 * (modules/bvh/include/lagrange/bvh/TriangleAABBTree.h:84-88; consumer modules/bvh/src/compute_mesh_distances.cpp:73), batched and
 * on the engine's own hierarchy. out_sqdist[i] = squared distance from q_i to the mesh, out_triangle[i] = a triangle that attains
 * it (input numbering), out_xyz[3i..] = the closest point on it; any output may be NULL (not all). max_distance > 0 bounds the
 * search: points farther away report max_distance^2, triangle -1 and their own position. flags: WN_QUERY_PRESORTED. */
WN_API wn_status wn_closest_point(const wn_engine* e, const float* q_xyz, int64_t n, float max_distance, uint32_t flags, float* out_sqdist,
                                  int32_t* out_triangle, float* out_xyz, void* stream);

/* ---- tree replication across GPUs ---------------------------------------------------------------------------
 * The packed tree is position independent: wn_tree_pack writes it into one contiguous buffer (host or device) that
 * can be broadcast (NCCL over NVLink) and adopted on another device with wn_create_from_packed. */
WN_API wn_status wn_tree_packed_size(const wn_engine* e, int64_t* nbytes);
WN_API wn_status wn_tree_pack(const wn_engine* e, void* dst, int64_t nbytes, void* stream);
WN_API wn_status wn_create_from_packed(const void* src, int64_t nbytes, const wn_options* opt, wn_engine** out);

/* Single-process multi-GPU: copies of `src` on the listed devices (device-to-device copies of the packed tree, all in
 * flight together; NVLink peer copies where the topology allows), out[i] on devices[i]. Destroy each with wn_destroy. */
WN_API wn_status wn_replicate(const wn_engine* src, const int32_t* devices, int32_t n_devices, wn_engine** out);
/* One lattice over several engines holding the same tree (the builder + its replicas): engine i evaluates rank i's share of the diagonal
 * sharding (wn_query_grid_sharded) on its own device from its own host thread, results are gathered in lattice order into HOST buffers (out_inside honours
 * WN_QUERY_OUT_BITS). No exchange step between devices. */
WN_API wn_status wn_query_grid_multi(wn_engine* const* engines, int32_t n_engines, const float origin[3], const float spacing[3],
                                     const int64_t dims[3], float beta, uint32_t flags, float* out_omega, uint8_t* out_inside);

/* ---- debug / parity hooks -----------------------------------------------------------------------------------
 * Raw per-node moments in the reference's stored form (23 floats, SURVEY.md A.4 order: P(3), maxPDist2, N(3),
 * Nxx,Nyy,Nzz, Nxy+Nyx, Nyz+Nzy, Nzx+Nxz, Nxxx,Nyyy,Nzzz, 2(Nxyz+Nyzx+Nzxy), 2Nxxy+Nyxx, 2Nxxz+Nzxx, 2Nyyz+Nzyy,
 * 2Nyyx+Nxyy, 2Nzzx+Nxzz, 2Nzzy+Nyzz) for hierarchy node `node` (internal nodes first, numbered as in the supplied
 * topology; then leaves: num_internal + triangle index). Host output. Needs keep_build_data = 1. */
WN_API wn_status wn_debug_node_moments(const wn_engine* e, int64_t first_node, int64_t count, float* out_23);
/* Hierarchy topology as built on the device: child[num_internal * width] with the wn_create_from_topology encoding. */
WN_API wn_status wn_debug_topology(const wn_engine* e, int32_t* child, int64_t capacity_nodes, int64_t* num_internal);
/* Class sizes of the tiles planned by the last tiled batch of this engine (K6'): out[4 * i + {0,1,2,3}] = conditional
 * records, direct records, gathered exact triangles, flags (1 = tile fell back to the generic traversal). Host output;
 * *num_tiles = tiles in that batch (call with out = NULL to size the buffer). */
WN_API wn_status wn_debug_last_plan(const wn_engine* e, int32_t* out, int64_t capacity_tiles, int64_t* num_tiles);
/* Stand-alone radix sort of (key,value) pairs on the device (K2), exposed for tests. Host pointers. */
WN_API wn_status wn_debug_sort_pairs_u64(uint64_t* keys, uint32_t* values, int64_t n, int32_t begin_bit, int32_t end_bit);
WN_API wn_status wn_debug_sort_pairs_u32(uint32_t* keys, uint32_t* values, int64_t n, int32_t begin_bit, int32_t end_bit);
/* FP32 FMA peak microbenchmark: runs `iters` dependent-chain FMAs per thread on a full grid, returns TFLOP/s. */
WN_API wn_status wn_debug_fma_peak(int32_t device, int32_t iters, float* tflops, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* WN_B200_H */
