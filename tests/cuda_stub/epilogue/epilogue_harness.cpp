// extern "C" doors into csrc/pmf_epilogue.cuh for tests/test_epilogue_cpu.py (host build, see cuda_runtime.h here).
#include "pmf_epilogue.cuh"

extern "C" {
void epi_noise_eval(int dist, int n, const float* z, const float* a, const float* th, float ord_eps, float margin,
                    float* l, float* g, int* observed) {
    for (int i = 0; i < n; ++i) {
        observed[i] = pmf::is_observed(a[i]) ? 1 : 0;
        l[i] = g[i] = 0.f;
        if (observed[i]) pmf::noise_eval(dist, z[i], a[i], th, ord_eps, margin, l[i], g[i]);
    }
}
void epi_threshold_grads(int dist, int n, const float* z, const float* a, const float* th, float ord_eps, float margin,
                         float* g1, float* g2) {
    for (int i = 0; i < n; ++i) {
        g1[i] = g2[i] = 0.f;
        if (pmf::is_observed(a[i])) pmf::noise_threshold_grads(dist, z[i], a[i], th, ord_eps, margin, g1[i], g2[i]);
    }
}
int epi_is_ordinal(int dist) { return pmf::is_ordinal(dist) ? 1 : 0; }
}
