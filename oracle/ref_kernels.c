/* CPU oracle, C restatement (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).
 *
 *  oracle_crc32_ieee    Go hash/crc32.ChecksumIEEE as used by stathat.com/c/consistent v1.0.0
 *                       (go.mod:25) through pkg/taskhandler/cluster.go:111,117.
 *  oracle_synth_fill    seeded weight generator, same integer hash as oracle/models.py.
 *  oracle_mlp_forward   y = act(x W + b) chained, the arithmetic the reference delegates to
 *                       TF-Serving (deploy/docker-compose/docker-compose.yaml:22-37): fp32 data,
 *                       accumulation in `double` when acc64 != 0 (arbiter) else in float.
 *  PARITY UNPINNED for numerics (no reference test holds a tensor); CRC pinned by the
 *  standard check value crc32("123456789") = 0xCBF43926.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

uint32_t oracle_crc32_ieee(const uint8_t* p, size_t n) {
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) {
        c ^= p[i];
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : (c >> 1);
    }
    return c ^ 0xFFFFFFFFu;
}

static uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

void oracle_synth_fill(float* dst, uint32_t seed, uint32_t tensor_id, uint64_t start, uint64_t n, float scale) {
    uint32_t k = mix32(seed * 0x9E3779B9u + tensor_id * 0x85EBCA6Bu + 0x165667B1u);
    for (uint64_t j = 0; j < n; ++j) {
        uint32_t h = mix32((uint32_t)(start + j) + k);
        float u = (float)(h >> 8) * (1.0f / 16777216.0f);
        dst[j] = (u * 2.0f - 1.0f) * scale;
    }
}

/* x:[B,dims[0]]  blob: W_l [in,out] row-major at w_off[l] (bytes), b_l at b_off[l]; relu[l]!=0 -> ReLU */
int oracle_mlp_forward(const float* blob, int n_layers, const int64_t* dims, const int64_t* w_off,
                       const int64_t* b_off, const int* relu, const float* x, int64_t B, float* y, int acc64) {
    int64_t maxd = 0;
    for (int l = 0; l <= n_layers; ++l) if (dims[l] > maxd) maxd = dims[l];
    float* cur = (float*)malloc(sizeof(float) * B * maxd);
    float* nxt = (float*)malloc(sizeof(float) * B * maxd);
    double* acc = (double*)malloc(sizeof(double) * maxd);
    if (!cur || !nxt || !acc) return -1;
    memcpy(cur, x, sizeof(float) * B * dims[0]);
    for (int l = 0; l < n_layers; ++l) {
        int64_t K = dims[l], N = dims[l + 1];
        const float* W = blob + w_off[l] / 4;
        const float* bias = blob + b_off[l] / 4;
        for (int64_t r = 0; r < B; ++r) {
            if (acc64) {
                for (int64_t n = 0; n < N; ++n) acc[n] = 0.0;
                for (int64_t k = 0; k < K; ++k) {
                    double xv = cur[r * K + k];
                    const float* wr = W + k * N;
                    for (int64_t n = 0; n < N; ++n) acc[n] += xv * (double)wr[n];
                }
                for (int64_t n = 0; n < N; ++n) {
                    float v = (float)(acc[n] + (double)bias[n]);
                    nxt[r * N + n] = (relu[l] && v < 0.f) ? 0.f : v;
                }
            } else {
                float* o = nxt + r * N;
                for (int64_t n = 0; n < N; ++n) o[n] = 0.f;
                for (int64_t k = 0; k < K; ++k) {
                    float xv = cur[r * K + k];
                    const float* wr = W + k * N;
                    for (int64_t n = 0; n < N; ++n) o[n] += xv * wr[n];
                }
                for (int64_t n = 0; n < N; ++n) {
                    float v = o[n] + bias[n];
                    o[n] = (relu[l] && v < 0.f) ? 0.f : v;
                }
            }
        }
        float* t = cur; cur = nxt; nxt = t;
    }
    memcpy(y, cur, sizeof(float) * B * dims[n_layers]);
    free(cur); free(nxt); free(acc);
    return 0;
}
