// Compile check and usage example of the C++ host mirror (needs a GPU to run):
//   g++ -std=c++17 integration/cpp/example.cpp -Lmantaray_b200 -lmantaray_b200 -Wl,-rpath,$PWD/mantaray_b200 -o example
#include <cstdio>
#include "many_rays.hpp"

int main()
{
    using namespace mantaray;
    ConstantSlope beach;                       // src/tests/linear_beach.rs shapes
    beach.h0 = 100.0f; beach.dhdx = -0.05f;
    ConstantCurrent still(0.0, 0.0);
    std::vector<RayState> rays = {{0, 0, 0.05 * std::cos(M_PI / 6), 0.05 * std::sin(M_PI / 6)}, {0, 0, 0.05, 0}};
    try {
        ManyRays waves(beach, still, rays);
        auto res = waves.trace_many(0.0, 1000.0, 1.0);
        for (auto &r : res) std::printf("%zu rows, last finite x = %.3f\n", r.t.size(), r.states[r.t.size() - 2].x);
    } catch (const Error &e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return e.code == MR_ERR_CUDA ? 0 : 1;  // no GPU on the build box: the library says so, loudly
    }
    return 0;
}
