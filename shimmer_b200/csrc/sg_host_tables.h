// Host-side table builders shared by libshimmer_gpu.so (sg_scene_create) and libshimmer_host.so (so that the CPU test suite can
// check them without a GPU).  Plain C++, no CUDA.
#pragma once
#include <cstdint>

namespace sg {

// Piecewise-linear spectra (spectrum.rs:295-440): `find_interval` (math.rs:322-333) returns the largest o in [0, n-2] with o == 0 or
// L[o] <= lambda.  The table holds that index for every integer wavelength LAMBDA_MIN..LAMBDA_MAX; the device starts from the entry
// of floor(lambda) and steps forward over the knots inside the same 1 nm bin (sg_shading.cuh spectrum_get), which ends on the
// identical interval.  Returns false -- keep the binary search -- when there is nothing to tabulate or the knots are not sorted
// (the binary search's answer is not "the last knot <= lambda" then).
static constexpr int kSpecLutMin = 360, kSpecLutMax = 830;
static constexpr int kSpecLutBins = kSpecLutMax - kSpecLutMin + 1;
inline bool build_spectrum_lut(const float* L, int n, uint16_t* out) {
    if (n < 2 || n > 65535) return false;
    for (int k = 1; k < n; ++k) if (!(L[k - 1] <= L[k])) return false;
    int o = 0;                                              // monotone in the wavelength
    for (int w = kSpecLutMin; w <= kSpecLutMax; ++w) {
        while (o + 1 <= n - 2 && L[o + 1] <= (float)w) ++o;
        out[w - kSpecLutMin] = (uint16_t)o;
    }
    return true;
}

}  // namespace sg
