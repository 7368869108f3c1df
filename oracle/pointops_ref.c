/*
 * ORACLE (test infrastructure, never shipped, never measured as the product).
 *
 * CPU restatement of the two live entry points of the third-party `pointops_cuda`
 * extension the reference calls (module git+https://github.com/Silverster98/pointops,
 * UNPINNED in /root/reference/requirements.txt:1, source absent from /root/reference):
 *
 *   furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx)
 *       call site: models/scene_models/pointops.py:10-27
 *   knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2)
 *       call site: models/scene_models/pointops.py:30-45
 *
 * Published algorithm restated (POSTECH point-transformer lib/pointops):
 *   FPS : per batch segment, first pick = first row of the segment; then repeat
 *         tmp[k] = min(tmp[k], |xyz[k]-xyz[last]|^2), next = argmax_k tmp[k].
 *   kNN : per query, scan every row of the SAME segment, keep the k smallest squared
 *         distances, output ascending; dist2 is squared (the python wrapper takes sqrt).
 *
 * PARITY UNPINNED at this boundary: the upstream source is not available offline, so tie
 * order (an artefact of its heap/reduction) is DEFINED here as lowest-index-wins, and the
 * CUDA kernels follow the same rule.  Arithmetic is fp32, d2 = (dx*dx + dy*dy) + dz*dz with
 * no FMA contraction (compile with -ffp-contract=off) so that the GPU kernels, which use
 * __fmul_rn/__fadd_rn in the same order, are bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline float sqdist(const float *a, const float *b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    float s = dx * dx;
    s = s + dy * dy;
    s = s + dz * dz;
    return s;
}

/* xyz [n,3]; offset[b], new_offset[b] cumulative ends; idx[m] out (global row indices). */
int oracle_fps(int b, const float *xyz, const int32_t *offset, const int32_t *new_offset,
               int32_t *idx) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < b; ++s) {
        int start_n = s == 0 ? 0 : offset[s - 1], end_n = offset[s];
        int start_m = s == 0 ? 0 : new_offset[s - 1], end_m = new_offset[s];
        int n = end_n - start_n;
        if (end_m <= start_m || n <= 0) continue;
        float *tmp = (float *)malloc(sizeof(float) * (size_t)n);
        for (int k = 0; k < n; ++k) tmp[k] = 1e10f; /* pointops.py:22 */
        int last = start_n;
        idx[start_m] = start_n;
        for (int j = start_m + 1; j < end_m; ++j) {
            float best = -1.0f;
            int besti = start_n;
            const float *pl = xyz + 3 * (size_t)last;
            for (int k = 0; k < n; ++k) {
                float d = sqdist(xyz + 3 * (size_t)(start_n + k), pl);
                float t = tmp[k] < d ? tmp[k] : d;
                tmp[k] = t;
                if (t > best) { best = t; besti = start_n + k; } /* strict > : lowest index wins */
            }
            idx[j] = besti;
            last = besti;
        }
        free(tmp);
    }
    return 0;
}

/* xyz [n,3], new_xyz [m,3]; idx [m,k], dist2 [m,k] out, ascending by (dist2, index). */
int oracle_knn(int b, int m, int k, const float *xyz, const float *new_xyz,
               const int32_t *offset, const int32_t *new_offset, int32_t *idx, float *dist2) {
    (void)m;
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < b; ++s) {
        int start_n = s == 0 ? 0 : offset[s - 1], end_n = offset[s];
        int start_m = s == 0 ? 0 : new_offset[s - 1], end_m = new_offset[s];
        for (int q = start_m; q < end_m; ++q) {
            float bd[64];
            int bi[64];
            int cnt = 0;
            const float *pq = new_xyz + 3 * (size_t)q;
            for (int j = start_n; j < end_n; ++j) {
                float d = sqdist(xyz + 3 * (size_t)j, pq);
                if (cnt < k) {
                    int p = cnt++;
                    while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
                    bd[p] = d; bi[p] = j;
                } else if (d < bd[k - 1]) { /* strict < : earlier (lower) index kept on ties */
                    int p = k - 1;
                    while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
                    bd[p] = d; bi[p] = j;
                }
            }
            for (int t = 0; t < k; ++t) {
                /* fewer than k candidates: upstream leaves the zero-initialised slots */
                idx[(size_t)q * k + t] = t < cnt ? bi[t] : 0;
                dist2[(size_t)q * k + t] = t < cnt ? bd[t] : 0.0f;
            }
        }
    }
    return 0;
}
