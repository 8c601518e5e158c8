// C++ test of the drop-in class, in the shape of the reference's own test file
// (adobe/lagrange modules/winding/tests/test_fast_winding_number.cpp): its "[!benchmark]" case draws 10 000 uniform
// samples in the mesh bbox with std::mt19937 and counts is_inside hits (:90-109). The reference asserts nothing; here
// the same protocol runs on a procedural torus (the dragon mesh is not available) and is checked against analytic
// inside/outside knowledge of the torus, plus the surface contract: error messages, move-only semantics, concurrent
// const queries (the OpenVDB/TBB call pattern, modules/volume/src/mesh_to_volume.cpp:175-183).
// Needs a GPU: run by tests/test_cpp_host_layer.py under the `gpu` marker.
#include <lagrange/winding/FastWindingNumber.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <memory>
#include <random>
#include <thread>
#include <vector>

using Scalar = float;
using Index = uint32_t;
using Mesh = lagrange::SurfaceMesh<Scalar, Index>;

static int g_failures = 0;
#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);     \
            ++g_failures;                                                   \
        }                                                                   \
    } while (0)

// torus around the Y axis, two triangles per quad, outward oriented
static Mesh make_torus(float R, float r, int nr, int np)
{
    Mesh mesh;
    const double two_pi = 6.283185307179586;
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < np; ++j) {
            const double u = two_pi * i / nr, w = two_pi * j / np;
            const double rad = R + r * std::cos(w);
            mesh.add_vertex({float(rad * std::cos(u)), float(r * std::sin(w)), float(rad * std::sin(u))});
        }
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < np; ++j) {
            const Index v00 = i * np + j, v01 = i * np + (j + 1) % np;
            const Index v10 = ((i + 1) % nr) * np + j, v11 = ((i + 1) % nr) * np + (j + 1) % np;
            mesh.add_triangle(v00, v01, v11);
            mesh.add_triangle(v00, v11, v10);
        }
    return mesh;
}

static bool torus_contains(float R, float r, const std::array<float, 3>& p, float margin)
{
    const double d = std::sqrt(double(p[0]) * p[0] + double(p[2]) * p[2]) - R;
    return std::sqrt(d * d + double(p[1]) * p[1]) < r - margin;
}
static bool torus_excludes(float R, float r, const std::array<float, 3>& p, float margin)
{
    const double d = std::sqrt(double(p[0]) * p[0] + double(p[2]) * p[2]) - R;
    return std::sqrt(d * d + double(p[1]) * p[1]) > r + margin;
}

int main()
{
    const float R = 5.f, r = 1.f;
    const Mesh mesh = make_torus(R, r, 200, 100);

    // --- surface contract ---------------------------------------------------------------------------------------
    {
        lagrange::SurfaceMesh<double, uint64_t> flat(2);
        flat.add_vertex({0.0, 0.0});
        bool threw = false;
        try {
            lagrange::winding::FastWindingNumber e(flat);
        } catch (const lagrange::Error& err) {
            threw = std::string(err.what()) == "Fast winding number engine only supports 3D meshes";
        }
        CHECK(threw);
        Mesh quad;
        for (int k = 0; k < 4; ++k) quad.add_vertex({float(k & 1), float(k >> 1), 0.f});
        quad.add_polygon({0, 1, 3, 2});
        threw = false;
        try {
            lagrange::winding::FastWindingNumber e(quad);
        } catch (const lagrange::Error& err) {
            threw = std::string(err.what()) == "Fast winding number engine only supports triangle meshes";
        }
        CHECK(threw);
        lagrange::winding::FastWindingNumber empty;
        threw = false;
        try {
            empty.is_inside({0.f, 0.f, 0.f});
        } catch (const lagrange::Error&) {
            threw = true;
        }
        CHECK(threw);
    }

    lagrange::winding::FastWindingNumber engine(mesh);
    std::printf("build %.3f ms, %lld triangles, tree %lld bytes\n", engine.build_milliseconds(), (long long)engine.num_triangles(),
                (long long)engine.tree_bytes());
    CHECK(engine.num_triangles() == 200 * 100 * 2);

    // --- known answers --------------------------------------------------------------------------------------------
    const float four_pi = 12.566370614359172f;
    CHECK(engine.is_inside({5.f, 0.f, 0.f}));
    CHECK(!engine.is_inside({0.f, 0.f, 0.f}));
    CHECK(!engine.is_inside({0.f, 3.f, 0.f}));
    CHECK(std::fabs(engine.solid_angle({5.f, 0.f, 0.f}) / four_pi - 1.f) < 1e-2f);
    CHECK(std::fabs(engine.solid_angle({0.f, 0.f, 0.f}) / four_pi) < 1e-2f);

    // move semantics: the moved-to engine answers, the moved-from one is empty
    lagrange::winding::FastWindingNumber moved(std::move(engine));
    CHECK(moved.is_inside({5.f, 0.f, 0.f}));
    engine = std::move(moved);
    CHECK(engine.is_inside({-5.f, 0.2f, 0.f}));

    // --- the reference's benchmark protocol: 10 000 uniform bbox samples, one is_inside call each -------------------
    const size_t num_samples = 10000;
    std::uniform_real_distribution<Scalar> px(-6.f, 6.f), py(-1.f, 1.f), pz(-6.f, 6.f);
    std::vector<std::array<float, 3>> samples(num_samples);
    {
        std::mt19937 gen;
        for (auto& s : samples) s = {px(gen), py(gen), pz(gen)};
    }
    size_t num_inside = 0, wrong = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (const auto& s : samples) {
        const bool in = engine.is_inside(s);
        num_inside += in;
        if (torus_contains(R, r, s, 0.02f) && !in) ++wrong;
        if (torus_excludes(R, r, s, 0.02f) && in) ++wrong;
    }
    auto t1 = std::chrono::steady_clock::now();
    const double single_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    std::printf("pimpl wrapper, %zu single-point calls: %.2f ms (%.2f us/call), %zu inside\n", num_samples, single_ms,
                1e3 * single_ms / num_samples, num_inside);
    CHECK(wrong == 0);

    // the batched overload must give the same answers as the single-point calls
    std::vector<uint8_t> batch(num_samples);
    t0 = std::chrono::steady_clock::now();
    engine.is_inside(samples[0].data(), num_samples, batch.data());
    t1 = std::chrono::steady_clock::now();
    size_t batch_inside = 0;
    for (auto b : batch) batch_inside += b;
    std::printf("batched overload, same %zu samples: %.3f ms, %zu inside\n", num_samples,
                std::chrono::duration<double, std::milli>(t1 - t0).count(), batch_inside);
    CHECK(batch_inside == num_inside);

    // tree vs exact mode on the same samples
    std::vector<float> om_tree(num_samples), om_exact(num_samples);
    engine.solid_angle(samples[0].data(), num_samples, om_tree.data());
    engine.exact_solid_angle(samples[0].data(), num_samples, om_exact.data());
    double max_err = 0;
    for (size_t i = 0; i < num_samples; ++i) max_err = std::max(max_err, double(std::fabs(om_tree[i] - om_exact[i])) / four_pi);
    std::printf("tree (beta=2) vs exact mode: max |dw| = %.3e\n", max_err);
    CHECK(max_err < 2e-2);

    // lattice overload agrees with explicit points
    {
        lagrange::winding::Lattice lat;
        lat.origin = {-6.5f, -1.5f, -6.5f};
        lat.spacing = {13.f / 40, 3.f / 16, 13.f / 40};
        lat.dims = {40, 16, 40};
        std::vector<uint8_t> g(40 * 16 * 40);
        engine.is_inside(lat, g.data());
        size_t mismatch = 0;
        for (int k = 0; k < 40; k += 3)
            for (int j = 0; j < 16; j += 2)
                for (int i = 0; i < 40; i += 3) {
                    const std::array<float, 3> p{lat.origin[0] + lat.spacing[0] * (float(i) + 0.5f), lat.origin[1] + lat.spacing[1] * (float(j) + 0.5f),
                                                 lat.origin[2] + lat.spacing[2] * (float(k) + 0.5f)};
                    mismatch += engine.is_inside(p) != (g[(size_t(k) * 16 + j) * 40 + i] != 0);
                }
        CHECK(mismatch == 0);

        // narrow-band signed distance on the same lattice: sign = is_inside, |d| <= band, torus of tube radius 1 around the
        // circle of radius 5 in the XZ plane => d ~ hypot(hypot(x, z) - 5, y) - 1 (the mesh is an inscribed polyhedron)
        const float band = 3.f * lat.spacing[0];
        std::vector<float> sdf(g.size());
        const int64_t active = engine.signed_distance(lat, band, sdf.data());
        size_t sign_mismatch = 0, in_band = 0;
        double worst = 0;
        for (int k = 0; k < 40; ++k)
            for (int j = 0; j < 16; ++j)
                for (int i = 0; i < 40; ++i) {
                    const size_t idx = (size_t(k) * 16 + j) * 40 + i;
                    sign_mismatch += (sdf[idx] < 0.f) != (g[idx] != 0);
                    CHECK(std::fabs(sdf[idx]) <= band);
                    if (std::fabs(sdf[idx]) < band) {
                        ++in_band;
                        const double x = lat.origin[0] + lat.spacing[0] * (i + 0.5), y = lat.origin[1] + lat.spacing[1] * (j + 0.5),
                                     z = lat.origin[2] + lat.spacing[2] * (k + 0.5);
                        worst = std::max(worst, std::fabs(double(sdf[idx]) - (std::hypot(std::hypot(x, z) - 5.0, y) - 1.0)));
                    }
                }
        std::printf("signed distance: %lld cells in the band, max deviation from the analytic torus %.3e\n", (long long)active, worst);
        CHECK(sign_mismatch == 0);
        CHECK(size_t(active) == in_band);
        CHECK(worst < 3e-2);

        // sparse narrow band == the dense block's active cells; closest point of the lattice's cell centres agrees with the band
        {
            const int64_t m = engine.signed_distance_sparse(lat, band, 0, nullptr, nullptr);
            CHECK(m == active);
            std::vector<int64_t> idx(m);
            std::vector<float> val(m);
            engine.signed_distance_sparse(lat, band, m, idx.data(), val.data());
            size_t bad = 0;
            for (int64_t k = 0; k < m; ++k) bad += val[k] != sdf[idx[k]] || (k > 0 && idx[k] <= idx[k - 1]);
            CHECK(bad == 0);
            std::vector<float> pts;
            for (int64_t k = 0; k < std::min<int64_t>(m, 2000); ++k) {
                const int64_t i = idx[k] % 40, j = (idx[k] / 40) % 16, kk = idx[k] / (40 * 16);
                pts.push_back(lat.origin[0] + lat.spacing[0] * (float(i) + 0.5f));
                pts.push_back(lat.origin[1] + lat.spacing[1] * (float(j) + 0.5f));
                pts.push_back(lat.origin[2] + lat.spacing[2] * (float(kk) + 0.5f));
            }
            const size_t np = pts.size() / 3;
            std::vector<float> sq(np), cp(3 * np);
            std::vector<int32_t> tri(np);
            engine.closest_point(pts.data(), np, sq.data(), tri.data(), cp.data());
            double worst_cp = 0;
            for (size_t k = 0; k < np; ++k) {
                worst_cp = std::max(worst_cp, std::fabs(std::sqrt(double(sq[k])) - std::fabs(double(val[k]))));
                CHECK(tri[k] >= 0 && tri[k] < 200 * 100 * 2);
            }
            std::printf("sparse band: %lld cells; closest point vs band distance: max deviation %.2e\n", (long long)m, worst_cp);
            CHECK(worst_cp < 1e-5);
        }

        // the balanced k-d hierarchy classifies the lattice like the LBVH does (different trees: allow the surface shell)
        lagrange::winding::FastWindingNumberOptions kd_opt;
        kd_opt.balanced_hierarchy = true;
        kd_opt.leaf_size = 4;
        lagrange::winding::FastWindingNumber kd(mesh, kd_opt);
        std::vector<uint8_t> g2(g.size());
        kd.is_inside(lat, g2.data());
        size_t differ = 0;
        for (size_t i = 0; i < g.size(); ++i) differ += g[i] != g2[i];
        CHECK(differ <= g.size() / 500);
    }

    // --- concurrent const queries from several host threads ---------------------------------------------------------
    {
        std::atomic<size_t> bad{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < 4; ++t)
            pool.emplace_back([&, t]() {
                for (size_t i = t; i < 2000; i += 4)
                    if (engine.is_inside(samples[i]) != (batch[i] != 0)) ++bad;
            });
        for (auto& th : pool) th.join();
        CHECK(bad == 0);
    }

    // --- the reference's production calling pattern: one is_inside per voxel from many worker threads (OpenVDB's TBB pool,
    //     modules/volume/src/mesh_to_volume.cpp:175-183). The single-point overloads walk a host copy of the packed tree: no
    //     launch, no lock. 16 threads x 10 000 calls; and the answers are those of the batched GPU traversal. ------------------
    {
        const int T = 16;
        const size_t per = 10000;
        std::vector<std::vector<std::array<float, 3>>> pts(T);
        for (int t = 0; t < T; ++t) {
            std::mt19937 gen(1234u + t);
            pts[t].resize(per);
            for (auto& s : pts[t]) s = {px(gen), py(gen), pz(gen)};
        }
        engine.is_inside(pts[0][0]); // first use copies the tree to the host
        std::vector<size_t> inside(T, 0);
        const auto a0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (int t = 0; t < T; ++t)
            pool.emplace_back([&, t]() {
                size_t c = 0;
                for (const auto& s : pts[t]) c += engine.is_inside(s);
                inside[t] = c;
            });
        for (auto& th : pool) th.join();
        const auto a1 = std::chrono::steady_clock::now();
        const double sec = std::chrono::duration<double>(a1 - a0).count();
        const double rate = double(T) * double(per) / sec;
        std::printf("single-point overload from %d host threads (%u hardware threads): %zu calls in %.2f ms = %.2f M calls/s aggregate\n", T,
                    std::thread::hardware_concurrency(), size_t(T) * per, 1e3 * sec, rate / 1e6);
        CHECK(rate > 1.0e6);
        // same points through the batched overload (GPU, per-point traversal semantics)
        size_t differ = 0, near_half = 0;
        double worst = 0;
        for (int t = 0; t < 2; ++t) {
            std::vector<uint8_t> b(per);
            std::vector<float> om(per);
            engine.is_inside(pts[t][0].data(), per, b.data());
            engine.solid_angle(pts[t][0].data(), per, om.data());
            for (size_t i = 0; i < per; ++i) {
                const float h = engine.solid_angle(pts[t][i]);
                worst = std::max(worst, double(std::fabs(h - om[i])) / four_pi);
                if (engine.is_inside(pts[t][i]) != (b[i] != 0)) {
                    ++differ;
                    near_half += std::fabs(om[i] / four_pi - 0.5f) < 1e-4f;
                }
            }
        }
        std::printf("host single-point vs batched GPU traversal: max |dw| = %.3e, is_inside differs on %zu points (%zu of them within 1e-4 of w = 1/2)\n",
                    worst, differ, near_half);
        CHECK(worst < 2e-5);
        CHECK(differ == near_half);
        // the launch-per-call variant still exists (options.host_single_point = false) and agrees
        lagrange::winding::FastWindingNumberOptions o2;
        o2.host_single_point = false;
        lagrange::winding::FastWindingNumber launch_engine(mesh, o2);
        size_t d2 = 0;
        for (size_t i = 0; i < 200; ++i) d2 += launch_engine.is_inside(pts[0][i]) != engine.is_inside(pts[0][i]);
        CHECK(d2 == 0);
        // bit-packed lattice overload
        lagrange::winding::Lattice lat;
        lat.origin = {-6.5f, -1.5f, -6.5f};
        lat.spacing = {13.f / 37, 3.f / 13, 13.f / 41};
        lat.dims = {37, 13, 41};
        const size_t n = 37 * 13 * 41;
        std::vector<uint8_t> bytes(n), bits((n + 7) / 8);
        engine.is_inside(lat, bytes.data());
        engine.is_inside_bits(lat, bits.data());
        size_t bitdiff = 0;
        for (size_t i = 0; i < n; ++i) bitdiff += ((bits[i >> 3] >> (i & 7)) & 1) != bytes[i];
        CHECK(bitdiff == 0);
    }

    // --- single-process multi-GPU through the drop-in class (options.devices); skipped on a one-GPU box -----------------------
    {
        lagrange::winding::FastWindingNumberOptions mo;
        mo.devices = {0, 1};
        bool have_two = true;
        std::unique_ptr<lagrange::winding::FastWindingNumber> multi;
        try {
            multi = std::make_unique<lagrange::winding::FastWindingNumber>(mesh, mo);
        } catch (const lagrange::Error& err) {
            have_two = false;
            std::printf("multi-GPU section skipped: %s\n", err.what());
        }
        if (have_two) {
            lagrange::winding::Lattice lat;
            lat.origin = {-6.5f, -1.5f, -6.5f};
            lat.spacing = {13.f / 96, 3.f / 40, 13.f / 96};
            lat.dims = {96, 40, 96};
            const size_t n = 96 * 40 * 96;
            std::vector<uint8_t> one(n), two(n), bits((n + 7) / 8);
            std::vector<float> omega(n);
            engine.is_inside(lat, one.data());
            engine.solid_angle(lat, omega.data());
            multi->is_inside(lat, two.data());
            multi->is_inside_bits(lat, bits.data());
            // if the tiled path is picked, a rank's planning blocks differ from the whole lattice's: answers may differ by the far-field
            // interpolation (<= 3e-5 * 4 pi), i.e. only inside the strict band |w - 1/2| <= 1e-3; the bits are the bytes of the same path
            size_t diff = 0, diff_outside_band = 0, bitdiff = 0;
            for (size_t i = 0; i < n; ++i) {
                const bool differs = one[i] != two[i];
                diff += differs;
                diff_outside_band += differs && std::fabs(omega[i] / (4.f * 3.14159265358979f) - 0.5f) > 1e-3f;
                bitdiff += ((bits[i >> 3] >> (i & 7)) & 1) != two[i];
            }
            std::printf("two GPUs, one process: %zu lattice points, %zu differ from the single-GPU answer (%zu outside the strict band)\n", n, diff,
                        diff_outside_band);
            CHECK(diff_outside_band == 0 && bitdiff == 0);
        }
    }

    std::printf(g_failures ? "FAILED (%d)\n" : "ALL PASSED\n", g_failures);
    return g_failures ? 1 : 0;
}
