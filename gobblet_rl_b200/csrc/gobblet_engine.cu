// gobblet_engine.cu -- kernels + C ABI (include/gobblet_b200.h) of the batched Gobblet engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a  (see build.py).  No CPU fallback exists.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gobblet_b200.h"
#include "gobblet_core.cuh"

namespace gbl {

#ifndef GBL_BLOCK
#define GBL_BLOCK 256   // threads per block (tuning knob; 256 measured best on B200)
#endif
#ifndef GBL_STEP_MIN_BLOCKS
#define GBL_STEP_MIN_BLOCKS (1024 / GBL_BLOCK)   // resident blocks per SM requested for step_kernel
#endif
#ifndef GBL_ROLLOUT_BLOCK
#define GBL_ROLLOUT_BLOCK GBL_BLOCK   // threads per block of the fused rollout kernel (dynamic shared memory)
#endif
constexpr int BLOCK = GBL_BLOCK, WARPS = BLOCK / 32;
constexpr int RBLOCK = GBL_ROLLOUT_BLOCK, RWARPS = RBLOCK / 32, RMIN_BLOCKS = 1024 / RBLOCK;

// ---- block-level statistics reduction: shuffles -> shared -> one atomic per slot per block ----
template <int NW = WARPS>
__device__ __forceinline__ void flush_stats(const Stats &st, bool valid, int64_t *stats) {
    __shared__ uint32_t red[NW][8];
    uint32_t v[8] = {st.episodes, st.p1w, st.p2w, st.steps, st.sumlen, st.illegal, st.both, st.maxlen};
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t x = valid ? v[i] : 0u;
        x = i == 7 ? __reduce_max_sync(0xFFFFFFFFu, x) : __reduce_add_sync(0xFFFFFFFFu, x);
        if (lane == 0) red[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        unsigned long long acc = 0;
        for (int w = 0; w < NW; ++w)
            acc = threadIdx.x == 7 ? max(acc, (unsigned long long)red[w][7]) : acc + red[w][threadIdx.x];
        if (acc) {
            if (threadIdx.x == 7) atomicMax(reinterpret_cast<long long *>(stats) + 7, (long long)acc);
            else atomicAdd(reinterpret_cast<unsigned long long *>(stats) + threadIdx.x, acc);
        }
    }
}

// ---- reset -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) reset_kernel(ulonglong2 *state, const uint8_t *which, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (g < n && (which == nullptr || which[g])) state[g] = make_ulonglong2(0ull, 0ull);
}

// ---- observe ---------------------------------------------------------------------------------------
template <bool kStreaming>
__global__ void __launch_bounds__(BLOCK)
observe_kernel(const ulonglong2 *__restrict__ state, int8_t *obs, int8_t *mask, uint8_t *agent_id, int64_t n) {
    __shared__ __align__(16) uint8_t stage[WARPS][STAGE_BYTES];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x, first = g - lane;
    if (first >= n) return;
    const bool valid = g < n;
    const int nvalid = (int)min((int64_t)32, n - first);
    stage_init(stage[warp], lane);
    Env e;
    env_clear(e);
    if (valid) env_unpack(e, state[g]);
    uint32_t u, up, m0, m1;
    occupancy(e, u, up);
    legal_mask(e.xo, e.yo, u, up, m0, m1);
    __syncwarp();
    stage_env(stage[warp], make_lane_cfg(lane), lane, e, m0, m1);
    __syncwarp();
    emit_chunk<kStreaming>(stage[warp], lane, obs + first * GBL_OBS_BYTES, mask + first * GBL_MASK_BYTES, nvalid);
    if (agent_id && valid) agent_id[g] = (uint8_t)e.agent;
    if (kBulkStore && lane == 0) bulk_store_wait_all();
}

// ---- externally driven step ----------------------------------------------------------------------
struct StepParams {
    ulonglong2 *state;
    const void *actions;
    int8_t *obs, *mask, *rew2;
    uint8_t *terminated, *truncated, *agent_id;
    int8_t *final_obs, *final_mask;
    int64_t *stats;
    int64_t n;
    uint32_t flags;
};

template <typename ActT, bool kStreaming>
__global__ void __launch_bounds__(BLOCK, GBL_STEP_MIN_BLOCKS) step_kernel(StepParams p) {
    __shared__ __align__(16) uint8_t stage[WARPS][STAGE_BYTES];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x, first = g - lane;
    const bool valid = g < p.n;
    Stats st = {0, 0, 0, 0, 0, 0, 0, 0};
    if (first < p.n) {
        const int nvalid = (int)min((int64_t)32, p.n - first);
        const LaneCfg cfg = make_lane_cfg(lane);
        stage_init(stage[warp], lane);
        Env e;
        env_clear(e);
        uint32_t action = 255u;
        if (valid) {
            env_unpack(e, p.state[g]);
            long long a = (long long)static_cast<const ActT *>(p.actions)[g];
            action = (a < 0 || a > 254) ? 255u : (uint32_t)a;
        }
        uint32_t u, up, m0, m1;
        occupancy(e, u, up);
        legal_mask(e.xo, e.yo, u, up, m0, m1);
        StepResult r = env_step<false>(e, m0, m1, action, p.flags, st);
        occupancy(e, u, up);
        legal_mask(e.xo, e.yo, u, up, m0, m1);
        const bool same_step = (p.flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_SAME_STEP;
        if (same_step && !r.acted) r.term = false;     // skipped env: nothing to reset
        if (same_step) {
            if (p.final_obs && p.final_mask) {      // terminal observation before it is replaced
                __syncwarp();
                stage_env(stage[warp], cfg, lane, e, m0, m1);
                __syncwarp();
                emit_chunk<kStreaming>(stage[warp], lane, p.final_obs + first * GBL_OBS_BYTES,
                                       p.final_mask + first * GBL_MASK_BYTES, nvalid);
            }
            if (r.term) {
                env_clear(e);
                occupancy(e, u, up);
                legal_mask(e.xo, e.yo, u, up, m0, m1);
            }
        }
        __syncwarp();
        stage_env(stage[warp], cfg, lane, e, m0, m1);
        __syncwarp();
        emit_chunk<kStreaming>(stage[warp], lane, p.obs + first * GBL_OBS_BYTES,
                               p.mask + first * GBL_MASK_BYTES, nvalid);
        if (kBulkStore && lane == 0) bulk_store_wait_all();
        if (valid) {
            p.state[g] = env_pack(e);
            if (p.rew2) *reinterpret_cast<char2 *>(p.rew2 + 2 * g) = make_char2((signed char)r.r1, (signed char)r.r2);
            if (p.terminated) p.terminated[g] = r.term;
            if (p.truncated) p.truncated[g] = r.trunc;
            if (p.agent_id) p.agent_id[g] = (uint8_t)e.agent;
        }
    }
    if (p.stats) flush_stats(st, valid, p.stats);
}

// ---- fused random-legal rollout -----------------------------------------------------------------
struct RolloutParams {
    ulonglong2 *state;
    int64_t n;
    int32_t T;
    uint64_t seed, env_id_base, step_base;
    const uint64_t *step_base_dev;
    int8_t *obs_out, *mask_out;
    int64_t obs_slot_stride, mask_slot_stride;
    int32_t ring;
    int8_t *rew_out;
    uint8_t *term_out, *agent_out, *action_log;
    int64_t *stats;
    uint32_t flags;
};

// kAux: any of rew_out / term_out / agent_out / action_log is requested (compiled out otherwise)
template <bool kFast, bool kStreaming, bool kAux>
__global__ void __launch_bounds__(RBLOCK, RMIN_BLOCKS) rollout_kernel(RolloutParams p) {
    extern __shared__ __align__(16) uint8_t stage_all[];      // RWARPS * STAGE_BYTES, one staging area per warp
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * RBLOCK + threadIdx.x, first = g - lane;
    uint8_t *const stage = stage_all + warp * STAGE_BYTES;
    const bool valid = g < p.n;
    Stats st = {0, 0, 0, 0, 0, 0, 0, 0};
    if (first < p.n) {
        const int nvalid = (int)min((int64_t)32, p.n - first);
        const LaneCfg cfg = make_lane_cfg(lane);
        const bool same_step = (p.flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_SAME_STEP;
        Env e;
        env_clear(e);
        if (valid) env_unpack(e, p.state[g]);
        const uint32_t plies_start = e.plies;
        uint32_t u, up, m0, m1;
        occupancy(e, u, up);
        legal_mask(e.xo, e.yo, u, up, m0, m1);
        stage_init(stage, lane);
        __syncwarp();
        const uint64_t step_base = p.step_base_dev ? *p.step_base_dev : p.step_base;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        uint32_t slot = (uint32_t)(step_base % (uint64_t)p.ring);
        const bool emit = p.obs_out != nullptr;
#pragma unroll 2
        for (int32_t t = 0; t < p.T; ++t) {     // two plies per trip: the own/opponent register swap becomes renaming
            const uint64_t s = step_base + (uint64_t)t;
            if (t == 0 || (s & 3u) == 0) rnd = draw_block(p.seed, p.env_id_base + (uint64_t)g, s, 0u);
            uint32_t action = 255u;
            if (kFast || !e.done) action = sample_action(m0, m1, pick_word(rnd, (uint32_t)s & 3u));
            StepResult r = env_step<kFast>(e, m0, m1, action, p.flags, st);
            if (r.term && same_step) env_clear(e);
            occupancy(e, u, up);
            legal_mask(e.xo, e.yo, u, up, m0, m1);
            if (kAux && valid) {
                const int64_t o = (int64_t)slot * p.n + g;
                if (p.rew_out) *reinterpret_cast<char2 *>(p.rew_out + 2 * o) = make_char2((signed char)r.r1, (signed char)r.r2);
                if (p.term_out) p.term_out[o] = r.term;
                if (p.agent_out) p.agent_out[o] = (uint8_t)e.agent;
                if (p.action_log) p.action_log[(int64_t)t * p.n + g] = r.acted ? (uint8_t)action : (uint8_t)255;
            }
            if (emit) {
                stage_env(stage, cfg, lane, e, m0, m1);
                __syncwarp();
                emit_chunk<kStreaming>(stage, lane, p.obs_out + (int64_t)slot * p.obs_slot_stride + first * GBL_OBS_BYTES,
                                       p.mask_out + (int64_t)slot * p.mask_slot_stride + first * GBL_MASK_BYTES, nvalid,
                                       (p.flags >> 8) & 3u);   // GBL_MEASURE_SKIP_*_STORES
                __syncwarp();
            }
            slot = slot + 1u == (uint32_t)p.ring ? 0u : slot + 1u;
        }
        if (kBulkStore && lane == 0) bulk_store_wait_all();
        if (valid) p.state[g] = env_pack(e);
        if (kFast) {   // every step was a live, legal step: these follow from the step count
            st.steps = (uint32_t)p.T;
            st.sumlen = plies_start + (uint32_t)p.T - e.plies;
            st.p2w = st.episodes - st.p1w;
        }
    }
    if (p.stats) flush_stats<RWARPS>(st, valid, p.stats);
}

// ---- masked-uniform sampler over int8 masks -----------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
sample_legal_kernel(const int8_t *__restrict__ mask, uint64_t seed, uint64_t env_id_base, uint64_t step,
                    const uint64_t *step_dev, int32_t *act, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (g >= n) return;
    if (step_dev) step = *step_dev;
    const int8_t *m = mask + g * GBL_MASK_BYTES;
    uint32_t m0 = 0, m1 = 0;
    for (int a = 0; a < 32; ++a) m0 |= (uint32_t)(m[a] != 0) << a;
    for (int a = 32; a < 54; ++a) m1 |= (uint32_t)(m[a] != 0) << (a - 32);
    uint4 b = draw_block(seed, env_id_base + (uint64_t)g, step, 0u);
    act[g] = (m0 | m1) ? (int32_t)sample_action(m0, m1, pick_word(b, (uint32_t)step & 3u)) : -1;
}

// ---- reference-layout views (board.py:33) -------------------------------------------------------
__global__ void __launch_bounds__(BLOCK)
export_squares_kernel(const ulonglong2 *__restrict__ state, int8_t *squares, uint8_t *agent, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (g >= n) return;
    ulonglong2 v = state[g];
    uint32_t x1 = (uint32_t)v.x & B27, y1 = (uint32_t)(v.x >> 27) & B27;
    uint32_t x2 = (uint32_t)v.y & B27, y2 = (uint32_t)(v.y >> 27) & B27;
    for (int i = 0; i < 27; ++i) {
        int l = i / 9, val = 0;
        if ((x1 >> i) & 1u) val = 2 * l + 1;
        if ((y1 >> i) & 1u) val = 2 * l + 2;
        if ((x2 >> i) & 1u) val = -(2 * l + 1);
        if ((y2 >> i) & 1u) val = -(2 * l + 2);
        squares[g * 27 + i] = (int8_t)val;
    }
    if (agent) agent[g] = (uint8_t)((v.x >> 54) & 1u);
}

__global__ void __launch_bounds__(BLOCK)
import_squares_kernel(ulonglong2 *state, const int8_t *__restrict__ squares, const uint8_t *__restrict__ agent, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (g >= n) return;
    uint32_t x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    for (int i = 0; i < 27; ++i) {
        int l = i / 9, val = squares[g * 27 + i];
        if (val == 2 * l + 1) x1 |= 1u << i;
        else if (val == 2 * l + 2) y1 |= 1u << i;
        else if (val == -(2 * l + 1)) x2 |= 1u << i;
        else if (val == -(2 * l + 2)) y2 |= 1u << i;
    }
    uint64_t meta = agent ? (uint64_t)(agent[g] & 1u) : 0ull;
    state[g] = make_ulonglong2((uint64_t)x1 | ((uint64_t)y1 << 27) | (meta << 54), (uint64_t)x2 | ((uint64_t)y2 << 27));
}

}  // namespace gbl

// =================================== C ABI ==========================================================
using namespace gbl;

static thread_local char g_err[512] = "";

static int fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
static int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return GBL_E_CUDA;
    }
    return 0;
}
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline unsigned grid_for(int64_t n) { return (unsigned)((n + BLOCK - 1) / BLOCK); }

extern "C" {

int gbl__set_error(const char *msg) { return fail(0, msg); }  // shared with gobblet_greedy.cu

int gbl_abi_version(void) { return GBL_ABI_VERSION; }
const char *gbl_last_error(void) { return g_err; }

int gbl_reset(void *state, int64_t n, void *stream) { return gbl_reset_masked(state, nullptr, n, stream); }

int gbl_reset_masked(void *state, const uint8_t *which, int64_t n, void *stream) {
    if (n < 0 || (n > 0 && (!state || !aligned16(state)))) return fail(GBL_E_INVALID, "gbl_reset: bad state pointer or n");
    if (n == 0) return 0;
    reset_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>((ulonglong2 *)state, which, n);
    return check_launch("gbl_reset");
}

int gbl_observe(const void *state, int8_t *obs, int8_t *mask, uint8_t *agent_id, int64_t n, void *stream) {
    if (n < 0) return fail(GBL_E_INVALID, "gbl_observe: n < 0");
    if (n == 0) return 0;
    if (!state || !obs || !mask || !aligned16(state) || !aligned16(obs) || !aligned16(mask))
        return fail(GBL_E_INVALID, "gbl_observe: state/obs/mask must be non-null and 16-byte aligned");
    observe_kernel<true><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>((const ulonglong2 *)state, obs, mask, agent_id, n);
    return check_launch("gbl_observe");
}

int gbl_step(void *state, const void *actions, int32_t action_bytes, int8_t *obs, int8_t *mask, int8_t *rew2,
             uint8_t *terminated, uint8_t *truncated, uint8_t *agent_id, int8_t *final_obs, int8_t *final_mask,
             int64_t *stats, int64_t n, uint32_t flags, void *stream) {
    if (n < 0) return fail(GBL_E_INVALID, "gbl_step: n < 0");
    if (n == 0) return 0;
    if (!state || !actions || !obs || !mask || !aligned16(state) || !aligned16(obs) || !aligned16(mask))
        return fail(GBL_E_INVALID, "gbl_step: state/actions/obs/mask must be non-null, state/obs/mask 16-byte aligned");
    if ((final_obs == nullptr) != (final_mask == nullptr) || (final_obs && (!aligned16(final_obs) || !aligned16(final_mask))))
        return fail(GBL_E_INVALID, "gbl_step: final_obs/final_mask must be given together, 16-byte aligned");
    if ((flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_MASK) return fail(GBL_E_INVALID, "gbl_step: bad autoreset mode");
    if (rew2 && (reinterpret_cast<uintptr_t>(rew2) & 1u)) return fail(GBL_E_INVALID, "gbl_step: rew2 must be 2-byte aligned");
    StepParams p = {(ulonglong2 *)state, actions, obs, mask, rew2, terminated, truncated, agent_id,
                    final_obs, final_mask, stats, n, flags};
    const bool plain = flags & GBL_STORE_DEFAULT_POLICY;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = grid_for(n);
#define GBL_LAUNCH_STEP(T)                                                        \
    do {                                                                          \
        if (plain) step_kernel<T, false><<<grid, BLOCK, 0, s>>>(p);               \
        else step_kernel<T, true><<<grid, BLOCK, 0, s>>>(p);                      \
    } while (0)
    switch (action_bytes) {
        case 1: GBL_LAUNCH_STEP(uint8_t); break;
        case 4: GBL_LAUNCH_STEP(int32_t); break;
        case 8: GBL_LAUNCH_STEP(int64_t); break;
        default: return fail(GBL_E_INVALID, "gbl_step: action_bytes must be 1, 4 or 8");
    }
#undef GBL_LAUNCH_STEP
    return check_launch("gbl_step");
}

int gbl_rollout_random(void *state, int64_t n, int32_t T, uint64_t seed, uint64_t env_id_base, uint64_t step_base,
                       const uint64_t *step_base_dev, int8_t *obs_out, int8_t *mask_out, int64_t obs_slot_stride, int64_t mask_slot_stride,
                       int32_t ring, int8_t *rew_out, uint8_t *term_out, uint8_t *agent_out, uint8_t *action_log,
                       int64_t *stats, uint32_t flags, void *stream) {
    if (n < 0 || T < 0) return fail(GBL_E_INVALID, "gbl_rollout_random: n or T < 0");
    if (n == 0 || T == 0) return 0;
    if (!state || !aligned16(state)) return fail(GBL_E_INVALID, "gbl_rollout_random: bad state pointer");
    if ((obs_out == nullptr) != (mask_out == nullptr)) return fail(GBL_E_INVALID, "gbl_rollout_random: obs_out and mask_out go together");
    if (ring < 1) return fail(GBL_E_INVALID, "gbl_rollout_random: ring must be >= 1");
    if (obs_out) {
        if (!aligned16(obs_out) || !aligned16(mask_out) || (obs_slot_stride & 15) || (mask_slot_stride & 15))
            return fail(GBL_E_INVALID, "gbl_rollout_random: obs/mask bases and slot strides must be multiples of 16 bytes");
        if (ring > 1 && (obs_slot_stride < n * GBL_OBS_BYTES || mask_slot_stride < n * GBL_MASK_BYTES))
            return fail(GBL_E_INVALID, "gbl_rollout_random: slot stride smaller than one step of output");
    }
    if ((flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_MASK) return fail(GBL_E_INVALID, "gbl_rollout_random: bad autoreset mode");
    if (rew_out && (reinterpret_cast<uintptr_t>(rew_out) & 1u)) return fail(GBL_E_INVALID, "gbl_rollout_random: rew_out must be 2-byte aligned");
    RolloutParams p = {(ulonglong2 *)state, n, T, seed, env_id_base, step_base, step_base_dev, obs_out, mask_out,
                       obs_slot_stride, mask_slot_stride, ring, rew_out, term_out, agent_out, action_log, stats, flags};
    // random legal actions never hit the illegal path; with same-step auto-reset no env is ever dead
    const bool fast = (flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_SAME_STEP;
    const bool plain = flags & GBL_STORE_DEFAULT_POLICY;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + RBLOCK - 1) / RBLOCK);
    const size_t smem = (size_t)RWARPS * STAGE_BYTES;
    const bool aux = rew_out || term_out || agent_out || action_log;
#define GBL_LAUNCH_ROLLOUT_1(F, S, A)                                                                            \
    do {                                                                                                         \
        if (smem > 48 * 1024)                                                                                    \
            cudaFuncSetAttribute(rollout_kernel<F, S, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        rollout_kernel<F, S, A><<<grid, RBLOCK, smem, s>>>(p);                                                   \
    } while (0)
#define GBL_LAUNCH_ROLLOUT(F, S)                                                    \
    do {                                                                            \
        if (aux) GBL_LAUNCH_ROLLOUT_1(F, S, true);                                  \
        else GBL_LAUNCH_ROLLOUT_1(F, S, false);                                     \
    } while (0)
    if (fast) {
        if (plain) GBL_LAUNCH_ROLLOUT(true, false);
        else GBL_LAUNCH_ROLLOUT(true, true);
    } else {
        if (plain) GBL_LAUNCH_ROLLOUT(false, false);
        else GBL_LAUNCH_ROLLOUT(false, true);
    }
#undef GBL_LAUNCH_ROLLOUT
#undef GBL_LAUNCH_ROLLOUT_1
    return check_launch("gbl_rollout_random");
}

int gbl_sample_legal(const int8_t *mask, uint64_t seed, uint64_t env_id_base, uint64_t step, const uint64_t *step_dev,
                     int32_t *act, int64_t n, void *stream) {
    if (n < 0) return fail(GBL_E_INVALID, "gbl_sample_legal: n < 0");
    if (n == 0) return 0;
    if (!mask || !act) return fail(GBL_E_INVALID, "gbl_sample_legal: null pointer");
    sample_legal_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(mask, seed, env_id_base, step, step_dev, act, n);
    return check_launch("gbl_sample_legal");
}

int gbl_export_squares(const void *state, int8_t *squares, uint8_t *agent, int64_t n, void *stream) {
    if (n < 0) return fail(GBL_E_INVALID, "gbl_export_squares: n < 0");
    if (n == 0) return 0;
    if (!state || !squares || !aligned16(state)) return fail(GBL_E_INVALID, "gbl_export_squares: bad pointer");
    export_squares_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>((const ulonglong2 *)state, squares, agent, n);
    return check_launch("gbl_export_squares");
}

int gbl_import_squares(void *state, const int8_t *squares, const uint8_t *agent, int64_t n, void *stream) {
    if (n < 0) return fail(GBL_E_INVALID, "gbl_import_squares: n < 0");
    if (n == 0) return 0;
    if (!state || !squares || !aligned16(state)) return fail(GBL_E_INVALID, "gbl_import_squares: bad pointer");
    import_squares_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>((ulonglong2 *)state, squares, agent, n);
    return check_launch("gbl_import_squares");
}

}  // extern "C"
