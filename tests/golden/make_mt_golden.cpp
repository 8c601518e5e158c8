// Generates tests/golden/mt19937_uniform_float.txt: what libstdc++ produces for the random stream of the reference's
// sample_points_in_mesh example (modules/winding/examples/sample_points_in_mesh.cpp:58-71): a default-seeded std::mt19937
// feeding three std::uniform_real_distribution<float> (x, y, z per point).   g++ -O2 make_mt_golden.cpp && ./a.out
#include <cstdio>
#include <random>
int main()
{
    const float lo[3] = {-6.3f, -1.05f, -6.3f}, hi[3] = {6.3f, 1.05f, 6.3f};
    std::uniform_real_distribution<float> px(lo[0], hi[0]), py(lo[1], hi[1]), pz(lo[2], hi[2]);
    std::mt19937 gen;
    printf("%.9g %.9g %.9g %.9g %.9g %.9g\n", lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    for (int k = 0; k < 4096; ++k) {
        const float x = px(gen);
        const float y = py(gen);
        const float z = pz(gen);
        printf("%.9g %.9g %.9g\n", x, y, z);
    }
    return 0;
}
