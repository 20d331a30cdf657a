// gobblet_engine.cu -- kernels + C ABI (include/gobblet_b200.h) of the batched Gobblet engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a  (see build.py).  No CPU fallback exists.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <type_traits>

#include "../../include/gobblet_b200.h"
#include "gobblet_core.cuh"

// gobblet_host.c (plain C, gcc): thread pool + bit -> byte expansion of the packed wire format
extern "C" {
void gblh_job_begin(const uint32_t *rec, int64_t n, int8_t *obs, int8_t *mask, int8_t *rew2, uint8_t *terminated,
                    uint8_t *truncated, uint8_t *agent_id, int32_t nthreads);
void gblh_job_publish(int64_t ready_envs);
int gblh_job_try_one(void);
void gblh_job_finish(void);
}

namespace gbl {

#ifndef GBL_BLOCK
#define GBL_BLOCK 256   // threads per block (tuning knob; 256 measured best on B200)
#endif
#ifndef GBL_STEP_MIN_BLOCKS
#define GBL_STEP_MIN_BLOCKS (1024 / GBL_BLOCK)   // resident blocks per SM requested for step_kernel
#endif
constexpr int BLOCK = GBL_BLOCK, WARPS = BLOCK / 32;

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
    // the block may retire once the copy engine has READ the image (the global writes complete by kernel end)
    if (kBulkStore && lane == 0) bulk_store_wait_read();
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
            // out-of-range actions are ILLEGAL moves (AssertOutOfBoundsWrapper territory, gobblet.py:115), mapped to
            // the sentinel 254; exactly 255 is reserved for "not stepped" under GBL_ACTION_SKIP_255
            long long a = (long long)static_cast<const ActT *>(p.actions)[g];
            action = (a < 0 || a > 255) ? 254u : (uint32_t)a;
        }
        uint32_t u, up, m0, m1;
        occupancy(e, u, up);
        legal_mask(e.xo, e.yo, u, up, m0, m1);
        StepResult r = env_step<false>(e, m0, m1, action, p.flags, st);
        occupancy(e, u, up);
        legal_mask(e.xo, e.yo, u, up, m0, m1);
        const bool same_step = (p.flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_SAME_STEP;
        // a skipped env is never reset; an env that ARRIVES finished (state imported from an "off" / "next_step"
        // run) reports its flags once more and is reset by this call, like the oracle's gbo_env_step
        const bool do_reset = same_step && r.term && !r.skipped;
        if (same_step) {
            if (p.final_obs && p.final_mask) {      // terminal observation before it is replaced
                __syncwarp();
                stage_env(stage[warp], cfg, lane, e, m0, m1);
                __syncwarp();
                emit_chunk<kStreaming>(stage[warp], lane, p.final_obs + first * GBL_OBS_BYTES,
                                       p.final_mask + first * GBL_MASK_BYTES, nvalid);
            }
            if (do_reset) {
                env_clear(e);
                occupancy(e, u, up);
                legal_mask(e.xo, e.yo, u, up, m0, m1);
            }
        }
        __syncwarp();
        if (same_step && p.final_obs && p.final_mask) stage_recycle(stage[warp], lane);
        stage_env(stage[warp], cfg, lane, e, m0, m1);
        __syncwarp();
        emit_chunk<kStreaming>(stage[warp], lane, p.obs + first * GBL_OBS_BYTES,
                               p.mask + first * GBL_MASK_BYTES, nvalid);
        if (valid) {
            p.state[g] = env_pack(e);
            if (p.rew2) *reinterpret_cast<char2 *>(p.rew2 + 2 * g) = make_char2((signed char)r.r1, (signed char)r.r2);
            if (p.terminated) p.terminated[g] = r.term;
            if (p.truncated) p.truncated[g] = r.trunc;
            if (p.agent_id) p.agent_id[g] = (uint8_t)e.agent;
        }
        if (kBulkStore && lane == 0) bulk_store_wait_read();   // shared memory may go once the engine has read the image
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
    int8_t *final_obs_out, *final_mask_out;
    int64_t *stats;
    uint32_t flags;
};

// kFast: same-step auto-reset, so every env is live at every step and every sampled action is legal (a valid
//        position always has >= 10 legal actions: both large pieces of the mover can move and <= 4 squares carry
//        a large top).  The one exception is an env that ARRIVES finished (state imported from an "off" /
//        "next_step" run): its first step only resets it, exactly as the general path and the oracle do.
// kAux:  any of rew_out / term_out / agent_out / action_log / final_*_out is requested (compiled out otherwise)
// kBulk: the observation image leaves through the copy engine (TMA bulk store); false = 32-lane LDS.128 -> STG.128 copy,
//        which has no asynchronous completion to wait for (small batches: one warp per scheduler, latency-bound)
// kPart:  PART_BOTH, or one of TWO warps that share the 32 envs of a chunk: both run the (register-only, deterministic)
//         game logic, one stages and emits the observations (and owns state, statistics and the aux outputs), the
//         other the masks.  For small batches (fewer warps than schedulers) a lockstep step is one warp's dependent
//         instruction chain; this takes the other stream's ~100 instructions out of it.
template <bool kFast, bool kStreaming, bool kAux, int kBlock, bool kBulk, int kPart>
__device__ __forceinline__ void rollout_body(const RolloutParams &p, const int64_t block) {
    extern __shared__ __align__(16) uint8_t stage_all[];      // (kBlock/32) * STAGE_BYTES, one staging area per warp
    constexpr bool kOwner = kPart != PART_MASK;               // stores state / statistics / aux outputs
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = block * kBlock + threadIdx.x, first = g - lane;
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
        uint32_t plies_start = e.plies, live_steps = (uint32_t)p.T;
        const bool dead0 = kFast && e.done;
        uint32_t u, up, m0, m1;
        occupancy(e, u, up);
        legal_mask(e.xo, e.yo, u, up, m0, m1);
        stage_init(stage, lane);
        __syncwarp();
        const uint64_t step_base = p.step_base_dev ? *p.step_base_dev : p.step_base;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        uint32_t slot = (p.flags & GBL_SLOT_FROM_ZERO) ? 0u : (uint32_t)(step_base % (uint64_t)p.ring);
        const bool emit = p.obs_out != nullptr;
        const uint32_t opts = ((p.flags >> 8) & 3u) |        // GBL_MEASURE_SKIP_*_STORES
                              (p.ring >= p.T ? 0u : p.ring == 1 ? EMIT_REUSE_ALWAYS : EMIT_REUSE_RING);
        const uint32_t initial = kAux && (p.flags & GBL_EMIT_INITIAL) ? 1u : 0u;
        if (initial) {                          // trajectory-buffer layout: slot 0 = the observation before the first step
            if (emit) {
                stage_recycle<kBulk && kBulkStore, kPart>(stage, lane);
                stage_env<kBulk && kBulkStore, kPart>(stage, cfg, lane, e, m0, m1);
                __syncwarp();
                emit_chunk<kStreaming, kBulk && kBulkStore, kPart>(stage, lane, p.obs_out + first * GBL_OBS_BYTES, p.mask_out + first * GBL_MASK_BYTES, nvalid, opts);
                __syncwarp();
            }
            if (kOwner && valid && p.agent_out) p.agent_out[g] = (uint8_t)e.agent;
            slot = 1u;
        }
        // byte offsets of this warp's chunk / this lane's element in the current slot, advanced by one slot per step
        // (a 64-bit add each instead of a 64-bit multiply-add per store address)
        int64_t obs_off = (int64_t)slot * p.obs_slot_stride + first * GBL_OBS_BYTES;
        int64_t mask_off = (int64_t)slot * p.mask_slot_stride + first * GBL_MASK_BYTES;
        int64_t aux_off = (int64_t)slot * p.n + g, log_off = g;
        const int64_t fin_obs_shift = (int64_t)initial * p.obs_slot_stride, fin_mask_shift = (int64_t)initial * p.mask_slot_stride;
        const int64_t aux_shift = (int64_t)initial * p.n;                     // [T+1] observation / agent slots vs [T] reward / flag slots
        // One lockstep step.  kFirst (t == 0, peeled): the Philox block is always drawn, and only here can a fast-path env
        // be one that ARRIVED finished -- the steady-state body carries neither test.  The body avoids divergent
        // regions (selects and predicated stores instead of branches) so that the emission of one step and the
        // register-only game logic around it are scheduled together.
        auto body = [&](const int32_t t, auto first_tag) {
            constexpr bool kFirst = decltype(first_tag)::value;
            const uint64_t s = step_base + (uint64_t)t;
            if (kFirst) {
                rnd = draw_block(p.seed, p.env_id_base + (uint64_t)g, s, 0u);
                for (uint32_t k = 0; k < ((uint32_t)s & 3u); ++k) rnd = make_uint4(rnd.y, rnd.z, rnd.w, rnd.x);
            } else if ((s & 3u) == 0) {
                rnd = draw_block(p.seed, p.env_id_base + (uint64_t)g, s, 0u);
            }
            const uint32_t draw = rnd.x;        // word (s & 3) of the block: the vector is rotated one word per step
            rnd = make_uint4(rnd.y, rnd.z, rnd.w, rnd.x);
            uint32_t action = 255u;
            StepResult r;
            if (kFirst && kFast && dead0) {     // reset-only step of an env that arrived finished
                r = {0, 0, true, e.trunc != 0, false, false};
                plies_start = 0; --live_steps;
            } else {
                if (kFast) action = sample_action<true>(m0, m1, draw);
                else if (!e.done) action = sample_action(m0, m1, draw);
                r = env_step<kFast>(e, m0, m1, action, p.flags, st);
            }
            occupancy(e, u, up);
            legal_mask(e.xo, e.yo, u, up, m0, m1);
            if (kAux && p.final_obs_out) {      // the observation a same-step reset is about to replace
                stage_recycle<kBulk && kBulkStore, kPart>(stage, lane);
                stage_env<kBulk && kBulkStore, kPart>(stage, cfg, lane, e, m0, m1);
                __syncwarp();
                emit_chunk<kStreaming, kBulk && kBulkStore, kPart>(stage, lane, p.final_obs_out + (obs_off - fin_obs_shift),
                                       p.final_mask_out + (mask_off - fin_mask_shift), nvalid, opts);
                __syncwarp();
            }
            {                                   // raw_env.reset (same-step): empty board, player_1 to move, every action legal
                const bool rs = r.term && same_step;
                e.xo = rs ? 0u : e.xo; e.yo = rs ? 0u : e.yo; e.xp = rs ? 0u : e.xp; e.yp = rs ? 0u : e.yp;
                e.agent = rs ? 0u : e.agent; e.plies = rs ? 0u : e.plies; e.done = rs ? 0u : e.done; e.trunc = rs ? 0u : e.trunc;
                m0 = rs ? 0xFFFFFFFFu : m0; m1 = rs ? 0x003FFFFFu : m1;
            }
            if (kAux && kOwner && valid) {
                const int64_t oa = aux_off - aux_shift;
                if (p.rew_out) *reinterpret_cast<char2 *>(p.rew_out + 2 * oa) = make_char2((signed char)r.r1, (signed char)r.r2);
                if (p.term_out) p.term_out[oa] = r.term;
                if (p.agent_out) p.agent_out[aux_off] = (uint8_t)e.agent;
                if (p.action_log) p.action_log[log_off] = r.acted ? (uint8_t)action : (uint8_t)255;
            }
            if (emit) {
                stage_recycle<kBulk && kBulkStore, kPart>(stage, lane);
                stage_env<kBulk && kBulkStore, kPart>(stage, cfg, lane, e, m0, m1);
                __syncwarp();
                emit_chunk<kStreaming, kBulk && kBulkStore, kPart>(stage, lane, p.obs_out + obs_off, p.mask_out + mask_off, nvalid, opts);
                __syncwarp();
            }
            log_off += p.n;
            if (slot + 1u == (uint32_t)p.ring) {     // uniform, and rare in a trajectory (ring >= T): back to slot 0
                slot = 0u;
                obs_off = first * GBL_OBS_BYTES; mask_off = first * GBL_MASK_BYTES; aux_off = g;
            } else {
                ++slot;
                obs_off += p.obs_slot_stride; mask_off += p.mask_slot_stride; aux_off += p.n;
            }
        };
        body(0, std::true_type{});
#pragma unroll 2
        for (int32_t t = 1; t < p.T; ++t) body(t, std::false_type{});   // two plies per trip: the own/opponent register swap becomes renaming
        if (kOwner && valid) p.state[g] = env_pack(e);
        if (kFast) {   // every step but a reset-only first one was a live, legal step: these follow from the step count
            st.steps = live_steps;
            st.sumlen = plies_start + live_steps - e.plies;
            st.p2w = st.episodes - st.p1w;
        }
        if (kBulk && kBulkStore && kOwner && lane == 0) bulk_store_wait_read();
    }
    if (kOwner && p.stats) flush_stats<kBlock / 32>(st, valid, p.stats);
}

// kSplit: blocks come in pairs -- even blocks emit the observations of chunk blockIdx.x / 2, odd blocks its masks
template <bool kFast, bool kStreaming, bool kAux, int kBlock, bool kBulk, bool kSplit>
__global__ void __launch_bounds__(kBlock, kBlock == 32 ? 16 : 1024 / kBlock) rollout_kernel(RolloutParams p) {
    if (kSplit) {
        if (blockIdx.x & 1u) rollout_body<kFast, kStreaming, kAux, kBlock, kBulk, PART_MASK>(p, (int64_t)(blockIdx.x >> 1));
        else rollout_body<kFast, kStreaming, kAux, kBlock, kBulk, PART_OBS>(p, (int64_t)(blockIdx.x >> 1));
    } else {
        rollout_body<kFast, kStreaming, kAux, kBlock, kBulk, PART_BOTH>(p, (int64_t)blockIdx.x);
    }
}

// ---- masked-uniform sampler over int8 masks -----------------------------------------------------
// A warp owns 32 consecutive mask rows = 1728 contiguous bytes = 108 x 16 bytes: the lanes load them as
// coalesced 128-bit vectors, squeeze every 16 bytes into 16 bits ("byte != 0") and park the 1728-bit stream in
// shared memory; each lane then cuts its own 54 bits out of the stream.  (One thread per row reading 54 single
// bytes at a 54-byte stride -- the first version -- touches every 32-byte sector 32 times.)
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w) {   // 4 bytes -> 4 bits, bit i = byte i != 0
    const uint32_t t = (((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u;
    return (((t >> 7) * 0x01020408u) >> 24) & 0xFu;
}

__global__ void __launch_bounds__(BLOCK)
sample_legal_kernel(const int8_t *__restrict__ mask, uint64_t seed, uint64_t env_id_base, uint64_t step,
                    const uint64_t *step_dev, int32_t *act, int64_t n) {
    __shared__ uint32_t bits[WARPS][56];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x, first = g - lane;
    if (first >= n) return;
    if (step_dev) step = *step_dev;
    const int8_t *rows = mask + first * GBL_MASK_BYTES;
    uint32_t m0 = 0, m1 = 0;
    if (n - first >= 32 && (reinterpret_cast<uintptr_t>(rows) & 15u) == 0) {
        uint16_t *hb = reinterpret_cast<uint16_t *>(bits[warp]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t q = lane + 32u * i;
            if (i < 3 || q < MASK_VEC) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(rows) + q);
                hb[q] = (uint16_t)(nonzero_nibble(v.x) | (nonzero_nibble(v.y) << 4) | (nonzero_nibble(v.z) << 8) |
                                   (nonzero_nibble(v.w) << 12));
            }
        }
        if (lane < 2) bits[warp][54 + lane] = 0;
        __syncwarp();
        const uint32_t o = 54u * lane, w = o >> 5, sh = o & 31u;
        const uint32_t a = bits[warp][w], b = bits[warp][w + 1], c = bits[warp][w + 2];
        m0 = __funnelshift_r(a, b, sh);
        m1 = __funnelshift_r(b, c, sh) & 0x003FFFFFu;
    } else if (g < n) {     // ragged last warp / unaligned base: byte loads
        const int8_t *m = mask + g * GBL_MASK_BYTES;
        for (int a = 0; a < 32; ++a) m0 |= (uint32_t)(m[a] != 0) << a;
        for (int a = 32; a < 54; ++a) m1 |= (uint32_t)(m[a] != 0) << (a - 32);
    }
    if (g >= n) return;
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
import_squares_kernel(ulonglong2 *state, const int8_t *__restrict__ squares, const uint8_t *__restrict__ agent,
                      int32_t *invalid_count, int64_t n) {
    int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (g >= n) return;
    uint32_t x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    bool bad = false;
    for (int i = 0; i < 27; ++i) {
        int l = i / 9, val = squares[g * 27 + i];
        if (val == 2 * l + 1) x1 |= 1u << i;
        else if (val == 2 * l + 2) y1 |= 1u << i;
        else if (val == -(2 * l + 1)) x2 |= 1u << i;
        else if (val == -(2 * l + 2)) y2 |= 1u << i;
        else if (val != 0) bad = true;               // a piece on a level that is not its size's, or |val| > 6
    }
    const uint32_t w[4] = {x1, y1, x2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int f = 0; f < 3; ++f) bad |= __popc(w[k] & (0x1FFu << (9 * f))) > 1;   // a piece placed twice (board.py:94-95)
    if (bad) {
        x1 = y1 = x2 = y2 = 0;
        if (invalid_count) atomicAdd(invalid_count, 1);
    }
    uint64_t meta = agent ? (uint64_t)(agent[g] & 1u) : 0ull;
    state[g] = make_ulonglong2((uint64_t)x1 | ((uint64_t)y1 << 27) | (meta << 54), (uint64_t)x2 | ((uint64_t)y2 << 27));
}

// ---- packed wire format (include/gobblet_b200.h): the step for host-side consumers ----------------------
// observation BITMAP, bit pos*13+c (gobblet.py:188-208), from the mover-relative boards
__device__ __forceinline__ void obs_bits(const Env &e, uint64_t &lo, uint64_t &hi) {
    lo = hi = 0;
    const uint32_t w[4] = {e.xo, e.yo, e.xp, e.yp};
    const int plane0[4] = {0, 1, 6, 7};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const uint32_t fld = w[k] & (0x1FFu << (9 * f));              // one-hot or empty
            const uint32_t bit = 13u * (bfind(fld) - 9u * f) + (uint32_t)(plane0[k] + 2 * f);
            const uint64_t one = fld ? 1ull : 0ull;
            lo |= (bit < 64u ? one : 0ull) << (bit & 63u);
            hi |= (bit >= 64u ? one : 0ull) << (bit & 63u);
        }
    if (e.agent) {                                   // plane 12: bits 12 + 13 p  (gobblet.py:199-206)
        lo |= 0x0008004002001000ull;                 // 12, 25, 38, 51
        hi |= 0x0010008004002001ull;                 // 64, 77, 90, 103, 116
    }
}

__device__ __forceinline__ void store_record(uint32_t *rec, int64_t g, const Env &e, uint32_t m0, uint32_t m1,
                                             int r1, int r2, bool term, bool trunc) {
    uint64_t lo, hi;
    obs_bits(e, lo, hi);
    const uint32_t fl = (uint32_t)(r1 + 1) | ((uint32_t)(r2 + 1) << 2) | ((uint32_t)term << 4) | ((uint32_t)trunc << 5) |
                        (e.agent << 6);
    uint2 *out = reinterpret_cast<uint2 *>(rec + 6 * g);               // 24-byte records: three 8-byte stores
    out[0] = make_uint2((uint32_t)lo, (uint32_t)(lo >> 32));
    out[1] = make_uint2((uint32_t)hi, (uint32_t)(hi >> 32) | (fl << 21));
    out[2] = make_uint2(m0, m1);
}

struct PackedParams {
    ulonglong2 *state;
    const void *actions;
    uint32_t *rec, *final_rec;
    int64_t *stats;
    int64_t n;
    uint32_t flags;
};

template <typename ActT>
__global__ void __launch_bounds__(BLOCK) step_packed_kernel(PackedParams p) {
    const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = g < p.n;
    Stats st = {0, 0, 0, 0, 0, 0, 0, 0};
    if (valid) {
        Env e;
        env_unpack(e, p.state[g]);
        long long a = (long long)static_cast<const ActT *>(p.actions)[g];
        const uint32_t action = (a < 0 || a > 255) ? 254u : (uint32_t)a;
        uint32_t u, up, m0, m1;
        occupancy(e, u, up);
        legal_mask(e.xo, e.yo, u, up, m0, m1);
        StepResult r = env_step<false>(e, m0, m1, action, p.flags, st);
        occupancy(e, u, up);
        legal_mask(e.xo, e.yo, u, up, m0, m1);
        const bool same_step = (p.flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_SAME_STEP;
        if (same_step) {
            if (p.final_rec) store_record(p.final_rec, g, e, m0, m1, r.r1, r.r2, r.term, r.trunc);
            if (r.term && !r.skipped) {
                env_clear(e);
                m0 = 0xFFFFFFFFu; m1 = 0x003FFFFFu;
            }
        }
        store_record(p.rec, g, e, m0, m1, r.r1, r.r2, r.term, r.trunc);
        p.state[g] = env_pack(e);
    }
    if (p.stats) flush_stats(st, valid, p.stats);
}

__global__ void __launch_bounds__(BLOCK) observe_packed_kernel(const ulonglong2 *__restrict__ state, uint32_t *rec, int64_t n) {
    const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (g >= n) return;
    Env e;
    env_unpack(e, state[g]);
    uint32_t u, up, m0, m1;
    occupancy(e, u, up);
    legal_mask(e.xo, e.yo, u, up, m0, m1);
    store_record(rec, g, e, m0, m1, 0, 0, e.done != 0, e.trunc != 0);
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

// threads per block of the fused rollout: 256 once the grid fills the machine; one-warp blocks for small batches
// so that every SM (x 4 schedulers) gets warps and the hardware spreads them evenly (BASELINE config 2, 4096
// envs, is latency-bound: ~1500 cycles of dependent instructions per warp-step).  Measured on a B200
// (benchmarks/sweep_small_batch.py, profiles/r2_sweep_small_batch.json): 32-thread blocks are fastest up to 64 Ki
// envs; below ~24 Ki envs copying the observation image with the lanes beats the copy engine (8 % at 16 Ki).
static int rollout_block_for(int64_t n, uint32_t flags) {
    const uint32_t hint = (flags >> GBL_BLOCK_HINT_SHIFT) & 7u;
    if (hint) return hint == 1 ? 32 : hint == 2 ? 64 : hint == 3 ? 128 : 256;
    const int64_t warps = (n + 31) / 32;
    if (warps <= 16 * 148) return 32;
    if (warps <= 32 * 148) return 128;
    return 256;
}

template <bool F, bool S, bool A, int B, bool K, bool SP = false>
static void launch_rollout(const RolloutParams &p, cudaStream_t s) {
    const unsigned grid = (unsigned)((p.n + B - 1) / B) * (SP ? 2u : 1u);
    const size_t smem = (size_t)(B / 32) * STAGE_BYTES;
    rollout_kernel<F, S, A, B, K, SP><<<grid, B, smem, s>>>(p);
}
template <bool F, bool S, bool A>
static void launch_rollout_b(const RolloutParams &p, int block, bool bulk, bool split, cudaStream_t s) {
    if constexpr (!S) {
        launch_rollout<F, S, A, 256, true>(p, s);                // plain stores: measurement aid, one size
    } else {
        switch (block) {
            case 32:
                if (split) return bulk ? launch_rollout<F, S, A, 32, true, true>(p, s) : launch_rollout<F, S, A, 32, false, true>(p, s);
                return bulk ? launch_rollout<F, S, A, 32, true>(p, s) : launch_rollout<F, S, A, 32, false>(p, s);
            case 64: return bulk ? launch_rollout<F, S, A, 64, true>(p, s) : launch_rollout<F, S, A, 64, false>(p, s);
            case 128: return launch_rollout<F, S, A, 128, true>(p, s);
            default: return launch_rollout<F, S, A, 256, true>(p, s);
        }
    }
}

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
                       int8_t *final_obs_out, int8_t *final_mask_out, int64_t *stats, uint32_t flags, void *stream) {
    if (n < 0 || T < 0) return fail(GBL_E_INVALID, "gbl_rollout_random: n or T < 0");
    if (n == 0 || T == 0) return 0;
    if (!state || !aligned16(state)) return fail(GBL_E_INVALID, "gbl_rollout_random: bad state pointer");
    if ((obs_out == nullptr) != (mask_out == nullptr)) return fail(GBL_E_INVALID, "gbl_rollout_random: obs_out and mask_out go together");
    if ((final_obs_out == nullptr) != (final_mask_out == nullptr) || (final_obs_out && !obs_out))
        return fail(GBL_E_INVALID, "gbl_rollout_random: final_obs_out and final_mask_out go together and need obs_out / mask_out");
    if (ring < 1) return fail(GBL_E_INVALID, "gbl_rollout_random: ring must be >= 1");
    if ((flags & GBL_EMIT_INITIAL) && (!(flags & GBL_SLOT_FROM_ZERO) || ring < T + 1))
        return fail(GBL_E_INVALID, "gbl_rollout_random: GBL_EMIT_INITIAL needs GBL_SLOT_FROM_ZERO and ring >= T + 1");
    if (obs_out) {
        if (!aligned16(obs_out) || !aligned16(mask_out) || (obs_slot_stride & 15) || (mask_slot_stride & 15))
            return fail(GBL_E_INVALID, "gbl_rollout_random: obs/mask bases and slot strides must be multiples of 16 bytes");
        if (final_obs_out && (!aligned16(final_obs_out) || !aligned16(final_mask_out)))
            return fail(GBL_E_INVALID, "gbl_rollout_random: final_obs_out / final_mask_out must be 16-byte aligned");
        if (ring > 1 && (obs_slot_stride < n * GBL_OBS_BYTES || mask_slot_stride < n * GBL_MASK_BYTES))
            return fail(GBL_E_INVALID, "gbl_rollout_random: slot stride smaller than one step of output");
    }
    if ((flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_MASK) return fail(GBL_E_INVALID, "gbl_rollout_random: bad autoreset mode");
    if (rew_out && (reinterpret_cast<uintptr_t>(rew_out) & 1u)) return fail(GBL_E_INVALID, "gbl_rollout_random: rew_out must be 2-byte aligned");
    RolloutParams p = {(ulonglong2 *)state, n, T, seed, env_id_base, step_base, step_base_dev, obs_out, mask_out,
                       obs_slot_stride, mask_slot_stride, ring, rew_out, term_out, agent_out, action_log,
                       final_obs_out, final_mask_out, stats, flags};
    // random legal actions never hit the illegal path; with same-step auto-reset no env stays dead
    const bool fast = (flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_SAME_STEP;
    const bool plain = flags & GBL_STORE_DEFAULT_POLICY;
    cudaStream_t s = (cudaStream_t)stream;
    const int block = rollout_block_for(n, flags);
    // lane copy instead of the copy engine: on request, or by default up to five warps per SM (latency-bound regime:
    // 0.65 vs 0.71 us per lockstep step at 16 Ki envs, equal at 32 Ki; profiles/r2_sweep_small_batch.json)
    const bool bulk = !((flags & GBL_NO_BULK_STORE_HINT) || (!((flags >> GBL_BLOCK_HINT_SHIFT) & 7u) && n <= 148 * 32 * 5));
    // two warps per chunk (observation warp / mask warp): on request, or by default while the doubled warps still find
    // a scheduler each (<= 9472 envs: 0.47 vs 0.59 us per lockstep step at 4096 envs, no gain left at 16 Ki), when
    // something is emitted at all
    const bool split = block == 32 && obs_out && !(flags & GBL_NO_SPLIT_HINT) &&
                       ((flags & GBL_SPLIT_HINT) || (!((flags >> GBL_BLOCK_HINT_SHIFT) & 7u) && n <= 148 * 32 * 2));
    const bool aux = rew_out || term_out || agent_out || action_log || final_obs_out || (flags & GBL_EMIT_INITIAL);
#define GBL_LAUNCH_ROLLOUT(F, S)                                                    \
    do {                                                                            \
        if (aux) launch_rollout_b<F, S, true>(p, block, bulk, split, s);            \
        else launch_rollout_b<F, S, false>(p, block, bulk, split, s);               \
    } while (0)
    if (fast) {
        if (plain) GBL_LAUNCH_ROLLOUT(true, false);
        else GBL_LAUNCH_ROLLOUT(true, true);
    } else {
        if (plain) GBL_LAUNCH_ROLLOUT(false, false);
        else GBL_LAUNCH_ROLLOUT(false, true);
    }
#undef GBL_LAUNCH_ROLLOUT
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

int gbl_import_squares(void *state, const int8_t *squares, const uint8_t *agent, int64_t n, int32_t *invalid_count,
                       void *stream) {
    if (n < 0) return fail(GBL_E_INVALID, "gbl_import_squares: n < 0");
    if (n == 0) return 0;
    if (!state || !squares || !aligned16(state)) return fail(GBL_E_INVALID, "gbl_import_squares: bad pointer");
    import_squares_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>((ulonglong2 *)state, squares, agent, invalid_count, n);
    return check_launch("gbl_import_squares");
}

int gbl_step_packed(void *state, const void *actions, int32_t action_bytes, uint32_t *rec, uint32_t *final_rec,
                    int64_t *stats, int64_t n, uint32_t flags, void *stream) {
    if (n < 0) return fail(GBL_E_INVALID, "gbl_step_packed: n < 0");
    if (n == 0) return 0;
    if (!state || !actions || !rec || !aligned16(state) || (reinterpret_cast<uintptr_t>(rec) & 7u) ||
        (reinterpret_cast<uintptr_t>(final_rec) & 7u))
        return fail(GBL_E_INVALID, "gbl_step_packed: state/actions/rec must be non-null, state 16-byte and rec 8-byte aligned");
    if ((flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_MASK) return fail(GBL_E_INVALID, "gbl_step_packed: bad autoreset mode");
    PackedParams p = {(ulonglong2 *)state, actions, rec, final_rec, stats, n, flags};
    cudaStream_t s = (cudaStream_t)stream;
    switch (action_bytes) {
        case 1: step_packed_kernel<uint8_t><<<grid_for(n), BLOCK, 0, s>>>(p); break;
        case 4: step_packed_kernel<int32_t><<<grid_for(n), BLOCK, 0, s>>>(p); break;
        case 8: step_packed_kernel<int64_t><<<grid_for(n), BLOCK, 0, s>>>(p); break;
        default: return fail(GBL_E_INVALID, "gbl_step_packed: action_bytes must be 1, 4 or 8");
    }
    return check_launch("gbl_step_packed");
}

int gbl_observe_packed(const void *state, uint32_t *rec, int64_t n, void *stream) {
    if (n < 0) return fail(GBL_E_INVALID, "gbl_observe_packed: n < 0");
    if (n == 0) return 0;
    if (!state || !rec || !aligned16(state) || (reinterpret_cast<uintptr_t>(rec) & 7u))
        return fail(GBL_E_INVALID, "gbl_observe_packed: bad pointer");
    observe_packed_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>((const ulonglong2 *)state, rec, n);
    return check_launch("gbl_observe_packed");
}

// ---- host side of the wire format: gobblet_host.c does the work, this file only knows CUDA events ------
int gbl_host_unpack_chunked(const uint32_t *rec, int64_t n, int32_t nchunks, const int64_t *chunk_end, void *const *events,
                            int8_t *obs, int8_t *mask, int8_t *rew2, uint8_t *terminated, uint8_t *truncated,
                            uint8_t *agent_id, int32_t nthreads) {
    if (n < 0 || nchunks < 0 || (n > 0 && (!rec || !obs || !mask))) return fail(GBL_E_INVALID, "gbl_host_unpack_chunked: bad argument");
    if (n == 0) return 0;
    if (nchunks > 0 && (!chunk_end || chunk_end[nchunks - 1] != n)) return fail(GBL_E_INVALID, "gbl_host_unpack_chunked: chunk_end must end at n");
    gblh_job_begin(rec, n, obs, mask, rew2, terminated, truncated, agent_id, nthreads);
    int rc = 0;
    for (int32_t c = 0; c < nchunks; ++c) {
        if (events && events[c]) {
            cudaError_t e = cudaEventSynchronize((cudaEvent_t)events[c]);
            if (e != cudaSuccess && rc == 0) {
                snprintf(g_err, sizeof(g_err), "gbl_host_unpack_chunked: %s", cudaGetErrorString(e));
                rc = GBL_E_CUDA;
            }
        }
        gblh_job_publish(chunk_end[c]);
    }
    if (nchunks == 0) gblh_job_publish(n);
    gblh_job_finish();
    return rc;
}

// measurement aid: wall-clock marks of the last gbl_step_host call of this thread (seconds since its entry):
// [0] enqueue done, [1 .. nchunks] chunk c published, [nchunks + 1] expansion finished
static thread_local double g_marks[66];
static thread_local int g_nmarks = 0;
static inline double now_s() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
int gbl_host_last_timing(double *out, int32_t cap) {
    const int k = g_nmarks < cap ? g_nmarks : cap;
    for (int i = 0; i < k; ++i) out[i] = g_marks[i];
    return g_nmarks;
}

int gbl_step_host(void *state, const uint8_t *actions_host, int64_t n, uint32_t flags, uint8_t *d_actions, uint32_t *d_rec,
                  uint32_t *h_rec, int32_t nchunks, const int64_t *chunk_end, void *stream, void *const *events,
                  int8_t *obs, int8_t *mask, int8_t *rew2, uint8_t *terminated, uint8_t *truncated, uint8_t *agent_id,
                  int64_t *stats, int32_t nthreads) {
    if (n < 0 || nchunks < 1) return fail(GBL_E_INVALID, "gbl_step_host: n < 0 or nchunks < 1");
    if (n == 0) return 0;
    if (!state || !actions_host || !d_actions || !d_rec || !h_rec || !chunk_end || !events || !aligned16(state) ||
        (reinterpret_cast<uintptr_t>(d_rec) & 7u) || (obs == nullptr) != (mask == nullptr))
        return fail(GBL_E_INVALID, "gbl_step_host: null / misaligned pointer, or obs and mask not given together");
    if (chunk_end[nchunks - 1] != n) return fail(GBL_E_INVALID, "gbl_step_host: chunk_end must end at n");
    for (int32_t c = 0; c < nchunks; ++c)
        if (chunk_end[c] < (c ? chunk_end[c - 1] : 0)) return fail(GBL_E_INVALID, "gbl_step_host: chunk_end must ascend");
    if ((flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_MASK) return fail(GBL_E_INVALID, "gbl_step_host: bad autoreset mode");
    cudaStream_t s = (cudaStream_t)stream;
    const bool expand = obs != nullptr;
    const double t_entry = now_s();
    const bool mark = nchunks <= 64;
    g_nmarks = 0;
    if (expand) gblh_job_begin(h_rec, n, obs, mask, rew2, terminated, truncated, agent_id, nthreads);   // workers spin on `ready`
    // one stream: actions H2D -> ONE step launch (packed records) -> the records D2H chunk by chunk, an event behind
    // each chunk.  While the later copies are still being enqueued, chunks that have already landed are published
    // to the expander (cudaEventQuery), so the host cores start on chunk 0 a few microseconds after it arrives.
    cudaMemcpyAsync(d_actions, actions_host, (size_t)n, cudaMemcpyHostToDevice, s);
    PackedParams p = {(ulonglong2 *)state, d_actions, d_rec, nullptr, stats, n, flags};
    step_packed_kernel<uint8_t><<<grid_for(n), BLOCK, 0, s>>>(p);
    int rc = check_launch("gbl_step_host");
    int32_t published = 0;
    int64_t a = 0;
    for (int32_t c = 0; c < nchunks && rc == 0; ++c) {
        const int64_t b = chunk_end[c];
        if (b > a) cudaMemcpyAsync(h_rec + 6 * a, d_rec + 6 * a, (size_t)(b - a) * 24, cudaMemcpyDeviceToHost, s);
        cudaEventRecord((cudaEvent_t)events[c], s);
        a = b;
        while (expand && published < c && cudaEventQuery((cudaEvent_t)events[published]) == cudaSuccess) {
            gblh_job_publish(chunk_end[published++]);
            if (mark) g_marks[published] = now_s() - t_entry;
        }
    }
    if (mark) g_marks[0] = now_s() - t_entry;
    static const int main_helps = getenv("GBL_HOST_MAIN_HELPS") ? atoi(getenv("GBL_HOST_MAIN_HELPS")) : 0;
    while (published < nchunks) {                    // the rest in order
        cudaError_t e = (expand && main_helps) ? cudaEventQuery((cudaEvent_t)events[published])
                                               : cudaEventSynchronize((cudaEvent_t)events[published]);
        if (e == cudaErrorNotReady) {                // (experiment) between two polls this thread expands too
            if (!gblh_job_try_one()) __builtin_ia32_pause();
            continue;
        }
        if (e != cudaSuccess && rc == 0) {
            snprintf(g_err, sizeof(g_err), "gbl_step_host: %s", cudaGetErrorString(e));
            rc = GBL_E_CUDA;
        }
        if (expand) gblh_job_publish(chunk_end[published]);
        ++published;
        if (mark) g_marks[published] = now_s() - t_entry;
    }
    if (expand) gblh_job_finish();
    if (mark) { g_marks[nchunks + 1] = now_s() - t_entry; g_nmarks = nchunks + 2; }
    return rc;
}

int gbl_host_unpack(const uint32_t *rec, int64_t n, int8_t *obs, int8_t *mask, int8_t *rew2, uint8_t *terminated,
                    uint8_t *truncated, uint8_t *agent_id, int32_t nthreads) {
    return gbl_host_unpack_chunked(rec, n, 0, nullptr, nullptr, obs, mask, rew2, terminated, truncated, agent_id, nthreads);
}

}  // extern "C"
