/* gobblet_host.c -- HOST side of the packed wire format (include/gobblet_b200.h): a persistent thread pool
 * that expands 24-byte records into the reference-shaped arrays a CPU consumer reads
 * (gobblet.py:179-215: observation int8 [3][3][13], action_mask int8 [54]; gobblet.py:255-263: rewards,
 * terminations; wrapper :110-117: truncations).
 *
 * Why it exists: the observation planes and the mask are 171 bytes of 0/1 per env.  Shipping them expanded
 * makes the end-to-end step PCIe-bound (176 B/env over a ~55 GB/s link); shipping BITS (24 B/env) and
 * expanding here moves the bound to host-memory write bandwidth, which the host cores reach together.
 * Plain C (gcc), no CUDA: gobblet_engine.cu wraps the job API below with the cudaEvent waits.
 *
 * Work distribution: a job is cut into blocks of GBLH_BLOCK envs handed out by an atomic counter; a block may
 * only be expanded once `ready` (published by the caller as D2H chunks complete) covers it.  Workers spin for
 * a while between jobs (steps arrive every millisecond or so) and then sleep on a condition variable.
 */
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GBLH_BLOCK 1024 /* envs per work item: a multiple of 64, so every item starts on a 64-byte boundary */
#define GBLH_MAX_THREADS 256
#define GBLH_SPIN 20000 /* pause-iterations a worker spins for a new job before it sleeps (~100 us) */

#if defined(__GNUC__)
#define GBL_API __attribute__((visibility("default")))
#else
#define GBL_API
#endif

typedef struct {
    const uint32_t *rec;
    int64_t n;
    int8_t *obs, *mask, *rew2;
    uint8_t *term, *trunc, *agent;
    /* fill job (measurement aid) when fill_dst != NULL */
    uint8_t *fill_dst;
    int64_t fill_bytes;
    int mode;
} gblh_job;

static struct {
    pthread_mutex_t api;   /* serialises jobs */
    pthread_mutex_t mu;    /* protects sleeping / generation */
    pthread_cond_t cv;
    pthread_t th[GBLH_MAX_THREADS];
    int nworkers;          /* threads created so far (excluding the caller) */
    int active;            /* workers taking part in the current job */
    _Atomic uint64_t generation;
    _Atomic int64_t next, ready;
    _Atomic int finished;
    _Atomic int stop;
    gblh_job job;
    int simd, store_mode;
    int init;
} P = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER};

/* ---- bit -> byte expansion ------------------------------------------------------------------------------- */
static uint64_t LUT8[256]; /* byte b -> 8 bytes, byte i = bit i of b */

static void lut_init(void) {
    for (int b = 0; b < 256; ++b) {
        uint64_t v = 0;
        for (int i = 0; i < 8; ++i) v |= (uint64_t)((b >> i) & 1) << (8 * i);
        LUT8[b] = v;
    }
}

static inline void flags_out(const gblh_job *j, int64_t i, uint32_t w3) {
    const uint32_t f = w3 >> 21;
    if (j->rew2) { j->rew2[2 * i] = (int8_t)((int)(f & 3u) - 1); j->rew2[2 * i + 1] = (int8_t)((int)((f >> 2) & 3u) - 1); }
    if (j->term) j->term[i] = (uint8_t)((f >> 4) & 1u);
    if (j->trunc) j->trunc[i] = (uint8_t)((f >> 5) & 1u);
    if (j->agent) j->agent[i] = (uint8_t)((f >> 6) & 1u);
}

/* portable path: one table lookup per 8 output bytes */
static void expand_table(const gblh_job *j, int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
        const uint32_t *r = j->rec + 6 * i;
        uint8_t b[24];
        memcpy(b, r, 24);
        int8_t *o = j->obs + 117 * i, *m = j->mask + 54 * i;
        for (int k = 0; k < 14; ++k) memcpy(o + 8 * k, &LUT8[b[k]], 8);          /* bits 0..111 */
        uint64_t t = LUT8[b[14] & 0x1Fu];                                        /* bits 112..116 */
        memcpy(o + 112, &t, 5);
        for (int k = 0; k < 6; ++k) memcpy(m + 8 * k, &LUT8[b[16 + k]], 8);      /* bits 0..47 */
        t = LUT8[b[22] & 0x3Fu];                                                 /* bits 48..53 */
        memcpy(m + 48, &t, 6);
        flags_out(j, i, r[3]);
    }
}

/* AVX-512BW: a 64-bit k-mask becomes 64 bytes of 0/1 in ONE instruction.  64 envs are expanded into an
 * L1-resident staging image (7488 + 3456 bytes, both multiples of 64) which then leaves with full-line
 * non-temporal stores: no read-for-ownership traffic on the 176 B/env the consumer will read later. */
/* rewards / flags of 16 consecutive envs at once: word 3 of their records (w3, collected by the caller) split into the
 * staging lines sf (rew2: 128 bytes at 0, terminated / truncated / agent_id: 64 bytes each at 128 / 192 / 256) */
__attribute__((target("avx512f,avx512bw"))) static inline void flags16_avx512(const uint32_t *w3, uint8_t *sf, int e) {
    const __m512i f = _mm512_srli_epi32(_mm512_load_si512((const void *)w3), 21);
    const __m512i m1 = _mm512_set1_epi32(1), m3 = _mm512_set1_epi32(3), mff = _mm512_set1_epi32(0xFF);
    const __m512i r1 = _mm512_and_si512(_mm512_sub_epi32(_mm512_and_si512(f, m3), m1), mff);
    const __m512i r2 = _mm512_and_si512(_mm512_sub_epi32(_mm512_and_si512(_mm512_srli_epi32(f, 2), m3), m1), mff);
    _mm256_store_si256((__m256i *)(sf + 2 * e), _mm512_cvtepi32_epi16(_mm512_or_si512(r1, _mm512_slli_epi32(r2, 8))));
    _mm_store_si128((__m128i *)(sf + 128 + e), _mm512_cvtepi32_epi8(_mm512_and_si512(_mm512_srli_epi32(f, 4), m1)));
    _mm_store_si128((__m128i *)(sf + 192 + e), _mm512_cvtepi32_epi8(_mm512_and_si512(_mm512_srli_epi32(f, 5), m1)));
    _mm_store_si128((__m128i *)(sf + 256 + e), _mm512_cvtepi32_epi8(_mm512_and_si512(_mm512_srli_epi32(f, 6), m1)));
}

__attribute__((target("avx512f,avx512bw"))) static void expand_avx512(const gblh_job *j, int64_t lo, int64_t hi, int direct) {
    const __m512i one = _mm512_set1_epi8(1);
    const uint64_t K53 = (1ull << 53) - 1ull, K54 = (1ull << 54) - 1ull;
    _Alignas(64) uint8_t sobs[64 * 117 + 64];
    _Alignas(64) uint8_t smask[64 * 54 + 64];
    _Alignas(64) uint8_t sf[320];
    _Alignas(64) uint32_t w3[64];
    int64_t i = lo;
    const int aligned = (((uintptr_t)(j->obs + 117 * lo) | (uintptr_t)(j->mask + 54 * lo) | (uintptr_t)(j->rew2 + 2 * lo) |
                          (uintptr_t)(j->term + lo) | (uintptr_t)(j->trunc + lo) | (uintptr_t)(j->agent + lo)) & 63u) == 0;
    for (; i + 64 <= hi && aligned && !direct; i += 64) {
        for (int e = 0; e < 64; ++e) {
            const uint32_t *r = j->rec + 6 * (i + e);
            const uint64_t k0 = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
            const uint64_t k1 = ((uint64_t)r[2] | ((uint64_t)r[3] << 32)) & K53;
            const uint64_t km = ((uint64_t)r[4] | ((uint64_t)r[5] << 32)) & K54;
            _mm512_storeu_si512((void *)(sobs + 117 * e), _mm512_maskz_mov_epi8((__mmask64)k0, one));
            _mm512_storeu_si512((void *)(sobs + 117 * e + 64), _mm512_maskz_mov_epi8((__mmask64)k1, one));  /* spills 11 zero bytes into env e+1 */
            _mm512_storeu_si512((void *)(smask + 54 * e), _mm512_maskz_mov_epi8((__mmask64)km, one));      /* spills 10 zero bytes */
            w3[e] = r[3];
        }
        for (int e = 0; e < 64; e += 16) flags16_avx512(w3 + e, sf, e);
        __m512i *od = (__m512i *)(j->obs + 117 * i), *md = (__m512i *)(j->mask + 54 * i);
        for (int l = 0; l < 117; ++l) _mm512_stream_si512(od + l, _mm512_load_si512((const void *)(sobs + 64 * l)));
        for (int l = 0; l < 54; ++l) _mm512_stream_si512(md + l, _mm512_load_si512((const void *)(smask + 64 * l)));
        if (j->rew2) {
            _mm512_stream_si512((__m512i *)(j->rew2 + 2 * i), _mm512_load_si512((const void *)sf));
            _mm512_stream_si512((__m512i *)(j->rew2 + 2 * i + 64), _mm512_load_si512((const void *)(sf + 64)));
        }
        if (j->term) _mm512_stream_si512((__m512i *)(j->term + i), _mm512_load_si512((const void *)(sf + 128)));
        if (j->trunc) _mm512_stream_si512((__m512i *)(j->trunc + i), _mm512_load_si512((const void *)(sf + 192)));
        if (j->agent) _mm512_stream_si512((__m512i *)(j->agent + i), _mm512_load_si512((const void *)(sf + 256)));
    }
    for (; i < hi; ++i) {   /* direct mode / unaligned base / tail: masked stores straight to the destination */
        const uint32_t *r = j->rec + 6 * i;
        const uint64_t k0 = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
        const uint64_t k1 = ((uint64_t)r[2] | ((uint64_t)r[3] << 32)) & K53;
        const uint64_t km = ((uint64_t)r[4] | ((uint64_t)r[5] << 32)) & K54;
        int8_t *o = j->obs + 117 * i, *m = j->mask + 54 * i;
        _mm512_storeu_si512((void *)o, _mm512_maskz_mov_epi8((__mmask64)k0, one));
        _mm512_mask_storeu_epi8((void *)(o + 64), (__mmask64)K53, _mm512_maskz_mov_epi8((__mmask64)k1, one));
        _mm512_mask_storeu_epi8((void *)m, (__mmask64)K54, _mm512_maskz_mov_epi8((__mmask64)km, one));
        flags_out(j, i, r[3]);
    }
    _mm_sfence();
}

__attribute__((target("avx512f"))) static void fill_nt(uint8_t *d, int64_t n) {
    const __m512i z = _mm512_set1_epi8(1);
    int64_t i = 0;
    for (; i < n && ((uintptr_t)(d + i) & 63u); ++i) d[i] = 1;
    for (; i + 64 <= n; i += 64) _mm512_stream_si512((__m512i *)(d + i), z);
    for (; i < n; ++i) d[i] = 1;
    _mm_sfence();
}

/* ---- the pool ----------------------------------------------------------------------------------------------- */
static void run_items(void) {
    const gblh_job *j = &P.job;
    if (j->fill_dst) {
        const int64_t item = 1 << 20, items = (j->fill_bytes + item - 1) / item;
        for (;;) {
            int64_t k = atomic_fetch_add_explicit(&P.next, 1, memory_order_relaxed);
            if (k >= items) break;
            int64_t a = k * item, b = a + item < j->fill_bytes ? a + item : j->fill_bytes;
            if (j->mode == 0 && P.simd) fill_nt(j->fill_dst + a, b - a);
            else memset(j->fill_dst + a, 1, (size_t)(b - a));
        }
        return;
    }
    const int64_t items = (j->n + GBLH_BLOCK - 1) / GBLH_BLOCK;
    for (;;) {
        int64_t k = atomic_fetch_add_explicit(&P.next, 1, memory_order_relaxed);
        if (k >= items) break;
        int64_t a = k * GBLH_BLOCK, b = a + GBLH_BLOCK < j->n ? a + GBLH_BLOCK : j->n;
        while (atomic_load_explicit(&P.ready, memory_order_acquire) < b) _mm_pause();   /* D2H of this block still in flight */
        if (P.simd) expand_avx512(j, a, b, P.store_mode);
        else expand_table(j, a, b);
    }
}

static void *worker(void *arg) {
    const int id = (int)(intptr_t)arg;
    uint64_t seen = 0;
    for (;;) {
        uint64_t g;
        int spins = 0;
        while ((g = atomic_load_explicit(&P.generation, memory_order_acquire)) == seen) {
            if (atomic_load_explicit(&P.stop, memory_order_relaxed)) return NULL;
            if (++spins < GBLH_SPIN) { _mm_pause(); continue; }
            pthread_mutex_lock(&P.mu);
            while (atomic_load_explicit(&P.generation, memory_order_acquire) == seen && !atomic_load(&P.stop))
                pthread_cond_wait(&P.cv, &P.mu);
            pthread_mutex_unlock(&P.mu);
            spins = 0;
        }
        seen = g;
        if (id < P.active) {
            run_items();
            atomic_fetch_add_explicit(&P.finished, 1, memory_order_release);
        }
    }
}

static int affinity_cores(void) {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        int c = CPU_COUNT(&set);
        if (c > 0) return c;
    }
    return 1;
}

static int resolve_threads(int nthreads) {
    int t = nthreads > 0 ? nthreads : affinity_cores();
    const char *env = getenv("GBL_HOST_THREADS");
    if (nthreads <= 0 && env && atoi(env) > 0) t = atoi(env);
    if (t > GBLH_MAX_THREADS) t = GBLH_MAX_THREADS;
    return t < 1 ? 1 : t;
}

static void pool_prepare(int nthreads) {   /* called with P.api held */
    if (!P.init) {
        lut_init();
        __builtin_cpu_init();
        P.simd = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f");
        const char *env = getenv("GBL_HOST_SIMD");
        if (env && atoi(env) == 0) P.simd = 0;
        env = getenv("GBL_HOST_STORE_MODE");
        if (env) P.store_mode = atoi(env) != 0;
        P.init = 1;
    }
    const int want = resolve_threads(nthreads) - 1;       /* the caller works too */
    while (P.nworkers < want) {
        if (pthread_create(&P.th[P.nworkers], NULL, worker, (void *)(intptr_t)P.nworkers) != 0) break;
        ++P.nworkers;
    }
    P.active = want < P.nworkers ? want : P.nworkers;
}

static void job_start(void) {
    atomic_store(&P.next, 0);
    atomic_store(&P.finished, 0);
    pthread_mutex_lock(&P.mu);
    atomic_fetch_add_explicit(&P.generation, 1, memory_order_release);
    pthread_cond_broadcast(&P.cv);
    pthread_mutex_unlock(&P.mu);
}

static void job_join(void) {
    run_items();
    while (atomic_load_explicit(&P.finished, memory_order_acquire) < P.active) _mm_pause();
}

/* job API used by gobblet_engine.cu (hidden visibility: not part of the C ABI) */
void gblh_job_begin(const uint32_t *rec, int64_t n, int8_t *obs, int8_t *mask, int8_t *rew2, uint8_t *terminated,
                    uint8_t *truncated, uint8_t *agent_id, int32_t nthreads) {
    pthread_mutex_lock(&P.api);
    pool_prepare(nthreads);
    gblh_job j = {rec, n, obs, mask, rew2, terminated, truncated, agent_id, NULL, 0, 0};
    P.job = j;
    atomic_store(&P.ready, 0);
    job_start();
}

void gblh_job_publish(int64_t ready_envs) { atomic_store_explicit(&P.ready, ready_envs, memory_order_release); }

/* The publishing thread helps between two event polls: expand ONE block, but only one that is already ready (it must
 * never wait for a block, since it is the thread that makes blocks ready).  Returns 1 if a block was expanded. */
int gblh_job_try_one(void) {
    const gblh_job *j = &P.job;
    int64_t k = atomic_load_explicit(&P.next, memory_order_relaxed);
    const int64_t items = (j->n + GBLH_BLOCK - 1) / GBLH_BLOCK;
    if (k >= items) return 0;
    const int64_t a = k * GBLH_BLOCK, b = a + GBLH_BLOCK < j->n ? a + GBLH_BLOCK : j->n;
    if (atomic_load_explicit(&P.ready, memory_order_acquire) < b) return 0;
    if (!atomic_compare_exchange_strong_explicit(&P.next, &k, k + 1, memory_order_relaxed, memory_order_relaxed)) return 0;
    if (P.simd) expand_avx512(j, a, b, P.store_mode);
    else expand_table(j, a, b);
    return 1;
}

void gblh_job_finish(void) {
    job_join();
    pthread_mutex_unlock(&P.api);
}

GBL_API int gbl_host_fill(void *dst, int64_t bytes, int32_t nthreads, int32_t mode) {
    if (!dst || bytes < 0) return -1;
    pthread_mutex_lock(&P.api);
    pool_prepare(nthreads);
    gblh_job j = {NULL, 0, NULL, NULL, NULL, NULL, NULL, NULL, (uint8_t *)dst, bytes, mode};
    P.job = j;
    job_start();
    job_join();
    pthread_mutex_unlock(&P.api);
    return 0;
}

GBL_API int gbl_host_threads(int32_t nthreads) { return resolve_threads(nthreads); }

GBL_API int gbl_host_simd(void) {
    pthread_mutex_lock(&P.api);
    pool_prepare(1);
    int s = P.simd;
    pthread_mutex_unlock(&P.api);
    return s;
}

GBL_API int gbl_host_set_store_mode(int32_t mode) {
    pthread_mutex_lock(&P.api);
    pool_prepare(1);
    P.store_mode = mode != 0;
    pthread_mutex_unlock(&P.api);
    return 0;
}
