/* CPU restatement (plain C) of the reference sampler and k-means assignment step.
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * oracle_fps_*           follows /root/reference/utils/utils.py:889-933 (see fps_oracle.py).
 * oracle_kmeans_assign_* follows the assignment step the reference delegates to
 *                        k_means_constrained (call sites data_proc/3_kmeans.py:78-82,
 *                        utils/utils.py:500-505): argmin_k ((x-c_k)**2).sum(-1), first minimum.
 * Build with -O2 -ffp-contract=off: the arithmetic must not be fused (NumPy does not fuse).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define DEFINE_FPS(NAME, T)                                                                   \
int NAME(const T* pc, int64_t n, int64_t row_stride, int64_t n_samples, int64_t start,        \
         int64_t* out) {                                                                      \
    if (n_samples > n || n_samples < 1 || start < 0 || start >= n) return -1;                 \
    T* d = (T*)malloc(sizeof(T) * (size_t)n);                                                 \
    if (!d) return -2;                                                                        \
    for (int64_t i = 0; i < n; ++i) {                                                         \
        const T* p = pc + i * row_stride;                                                     \
        if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) { free(d); return -3; }    \
        d[i] = (T)INFINITY;                                                                   \
    }                                                                                         \
    int64_t last = start;                                                                     \
    out[0] = last;                                                                            \
    d[last] = (T)-1;                                                                          \
    for (int64_t s = 1; s < n_samples; ++s) {                                                 \
        const T lx = pc[last * row_stride], ly = pc[last * row_stride + 1],                   \
                lz = pc[last * row_stride + 2];                                               \
        T best = (T)-1; int64_t bi = -1;                                                      \
        for (int64_t i = 0; i < n; ++i) {                                                     \
            T cur = d[i];                                                                     \
            if (cur < 0) continue;                    /* already picked: left the set */      \
            const T* p = pc + i * row_stride;                                                 \
            volatile T dx = lx - p[0], dy = ly - p[1], dz = lz - p[2];                        \
            volatile T xx = dx * dx, yy = dy * dy, zz = dz * dz;                              \
            volatile T xy = xx + yy;                                                          \
            T dist = xy + zz;                                                                 \
            if (dist < cur) { cur = dist; d[i] = cur; }                                       \
            if (cur > best) { best = cur; bi = i; }   /* strict: lowest index wins ties */    \
        }                                                                                     \
        last = bi;                                                                            \
        out[s] = last;                                                                        \
        d[last] = (T)-1;                                                                      \
    }                                                                                         \
    free(d);                                                                                  \
    return 0;                                                                                 \
}

DEFINE_FPS(oracle_fps_f32, float)
DEFINE_FPS(oracle_fps_f64, double)

int oracle_kmeans_assign_f32(const float* x, const float* c, int64_t n, int32_t k,
                             int32_t* labels, float* min_d2) {
    if (k < 1) return -1;
    for (int64_t i = 0; i < n; ++i) {
        float best = INFINITY; int32_t bj = 0;
        for (int32_t j = 0; j < k; ++j) {
            volatile float d0 = x[3 * i] - c[3 * j], d1 = x[3 * i + 1] - c[3 * j + 1],
                           d2 = x[3 * i + 2] - c[3 * j + 2];
            volatile float a = d0 * d0, b = d1 * d1, e = d2 * d2;
            volatile float ab = a + b;
            float dist = ab + e;
            if (dist < best) { best = dist; bj = j; }  /* strict: first minimum wins */
        }
        labels[i] = bj;
        if (min_d2) min_d2[i] = best;
    }
    return 0;
}
