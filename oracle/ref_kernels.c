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

/* Multi-threaded forward for the CPU baseline (bench.py cpu_baseline / --impl reference): the
 * arithmetic TF-Serving's intra-op thread pool would do for a batch-1 request, written the way a
 * bandwidth-bound GEMV wants it on a many-core host: K is split across threads (each streams a
 * contiguous block of W rows), private fp32 accumulators, then a reduction. Persistent pthread
 * pool (no OpenMP runtime in this image's gcc). */
#include <pthread.h>

typedef struct {
    const float* W; const float* bias; const float* cur; float* nxt; float* part;
    int64_t K, N, B; int relu; int T;
} mt_job;

static struct {
    int T; int started; int quit;
    pthread_t* th; pthread_barrier_t go, mid, done; mt_job job;
} g_pool;

static void mt_phase1(const mt_job* j, int t) {
    const int64_t k0 = j->K * t / j->T, k1 = j->K * (t + 1) / j->T;
    float* acc = j->part + (size_t)t * j->B * j->N;
    for (int64_t i = 0; i < j->B * j->N; ++i) acc[i] = 0.f;
    for (int64_t k = k0; k < k1; ++k) {
        const float* wr = j->W + k * j->N;
        for (int64_t r = 0; r < j->B; ++r) {
            const float xv = j->cur[r * j->K + k];
            float* a = acc + r * j->N;
            for (int64_t n = 0; n < j->N; ++n) a[n] += xv * wr[n];
        }
    }
}
static void mt_phase2(const mt_job* j, int t) {
    const int64_t tot = j->B * j->N, i0 = tot * t / j->T, i1 = tot * (t + 1) / j->T;
    for (int64_t i = i0; i < i1; ++i) {
        float s = 0.f;
        for (int tt = 0; tt < j->T; ++tt) s += j->part[(size_t)tt * tot + i];
        s += j->bias[i % j->N];
        j->nxt[i] = (j->relu && s < 0.f) ? 0.f : s;
    }
}
static void* mt_worker(void* arg) {
    const int t = (int)(intptr_t)arg;
    for (;;) {
        pthread_barrier_wait(&g_pool.go);
        if (g_pool.quit) return 0;
        mt_phase1(&g_pool.job, t);
        pthread_barrier_wait(&g_pool.mid);
        mt_phase2(&g_pool.job, t);
        pthread_barrier_wait(&g_pool.done);
    }
}
static int mt_start(int T) {
    if (g_pool.started && g_pool.T == T) return 0;
    if (g_pool.started) return -2;  /* one pool size per process */
    g_pool.T = T;
    g_pool.th = (pthread_t*)malloc(sizeof(pthread_t) * T);
    pthread_barrier_init(&g_pool.go, 0, T); pthread_barrier_init(&g_pool.mid, 0, T); pthread_barrier_init(&g_pool.done, 0, T);
    for (int t = 1; t < T; ++t) pthread_create(&g_pool.th[t], 0, mt_worker, (void*)(intptr_t)t);
    g_pool.started = 1;
    return 0;
}

int oracle_mlp_forward_mt(const float* blob, int n_layers, const int64_t* dims, const int64_t* w_off,
                          const int64_t* b_off, const int* relu, const float* x, int64_t B, float* y, int nthreads) {
    int64_t maxd = 0;
    for (int l = 0; l <= n_layers; ++l) if (dims[l] > maxd) maxd = dims[l];
    if (nthreads < 1) nthreads = 1;
    if (mt_start(nthreads)) return -2;
    float* cur = (float*)malloc(sizeof(float) * B * maxd);
    float* nxt = (float*)malloc(sizeof(float) * B * maxd);
    float* part = (float*)malloc(sizeof(float) * (size_t)nthreads * B * maxd);
    if (!cur || !nxt || !part) return -1;
    memcpy(cur, x, sizeof(float) * B * dims[0]);
    for (int l = 0; l < n_layers; ++l) {
        mt_job* j = &g_pool.job;
        j->W = blob + w_off[l] / 4; j->bias = blob + b_off[l] / 4; j->cur = cur; j->nxt = nxt; j->part = part;
        j->K = dims[l]; j->N = dims[l + 1]; j->B = B; j->relu = relu[l]; j->T = nthreads;
        pthread_barrier_wait(&g_pool.go);   /* thread 0 = the caller */
        mt_phase1(j, 0);
        pthread_barrier_wait(&g_pool.mid);
        mt_phase2(j, 0);
        pthread_barrier_wait(&g_pool.done);
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    memcpy(y, cur, sizeof(float) * B * dims[n_layers]);
    free(cur); free(nxt); free(part);
    return 0;
}
