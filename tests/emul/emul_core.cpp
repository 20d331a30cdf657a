// Host emulation of the rollout / step kernel bodies of gobblet_engine.cu, built from the SAME
// gobblet_core.cuh device functions (TEST ONLY).  One "warp" = 32 lanes run in lockstep; the
// one shuffle of stage_env is emulated with a record/replay pass.
#include "cuda_shim.h"
#include "../../include/gobblet_b200.h"
#include "../../gobblet_rl_b200/csrc/gobblet_core.cuh"
#include <string.h>
#include <vector>

thread_local ShflCtx g_shfl;
using namespace gbl;

// one persistent staging area per emulated warp, like the __shared__ array of the kernels
alignas(16) static uint8_t g_stage[STAGE_BYTES];
static bool g_stage_ready = false;

static void warp_emit(Env *e, uint32_t *m0, uint32_t *m1, int8_t *obs_chunk, int8_t *mask_chunk, int nvalid) {
    if (!g_stage_ready) {
        memset(g_stage, 0xAB, sizeof(g_stage));
        for (int lane = 0; lane < 32; ++lane) stage_init(g_stage, lane);
        g_stage_ready = true;
    }
    for (int pass = 0; pass < 2; ++pass)
        for (int lane = 0; lane < 32; ++lane) {
            g_shfl.pass = pass; g_shfl.call = 0; g_shfl.lane = lane;
            stage_env(g_stage, make_lane_cfg(lane), lane, e[lane], m0[lane], m1[lane]);
        }
    for (int lane = 0; lane < 32; ++lane) emit_chunk<true>(g_stage, lane, obs_chunk, mask_chunk, nvalid);
}

extern "C" __attribute__((visibility("default")))
void emul_rollout(unsigned long long *state, int64_t n, int32_t T, uint64_t seed, uint64_t env_id_base,
                  uint64_t step_base, uint32_t flags, int8_t *obs, int8_t *mask, int8_t *rew, uint8_t *term,
                  uint8_t *agent, uint8_t *log, int64_t *stats) {
    const bool same_step = (flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_SAME_STEP;
    const bool fast = same_step;
    for (int64_t first = 0; first < n; first += 32) {
        int nvalid = (int)std::min<int64_t>(32, n - first);
        Env e[32]; uint32_t m0[32], m1[32], plies_start[32], live_steps[32]; uint4 rnd[32]; Stats st[32]; bool dead0[32];
        for (int l = 0; l < 32; ++l) {
            env_clear(e[l]); memset(&st[l], 0, sizeof(Stats));
            if (l < nvalid) env_unpack(e[l], {state[2 * (first + l)], state[2 * (first + l) + 1]});
            plies_start[l] = e[l].plies; live_steps[l] = (uint32_t)T; dead0[l] = fast && e[l].done;
            uint32_t u, up; occupancy(e[l], u, up); legal_mask(e[l].xo, e[l].yo, u, up, m0[l], m1[l]);
        }
        for (int32_t t = 0; t < T; ++t) {
            uint64_t s = step_base + t;
            for (int l = 0; l < 32; ++l) {
                int64_t g = first + l;
                if (t == 0 || (s & 3) == 0) rnd[l] = draw_block(seed, env_id_base + g, s, 0);
                uint32_t action = 255;
                StepResult r;
                if (fast && dead0[l]) {
                    r = {0, 0, true, e[l].trunc != 0, false, false};
                    dead0[l] = false; plies_start[l] = 0; --live_steps[l];
                } else {
                    if (fast || !e[l].done) action = sample_action(m0[l], m1[l], pick_word(rnd[l], s & 3));
                    r = fast ? env_step<true>(e[l], m0[l], m1[l], action, flags, st[l])
                             : env_step<false>(e[l], m0[l], m1[l], action, flags, st[l]);
                }
                uint32_t u, up; occupancy(e[l], u, up); legal_mask(e[l].xo, e[l].yo, u, up, m0[l], m1[l]);
                if (r.term && same_step) { env_clear(e[l]); m0[l] = 0xFFFFFFFFu; m1[l] = 0x003FFFFFu; }
                if (l < nvalid) {
                    int64_t o = (int64_t)t * n + g;
                    rew[2 * o] = r.r1; rew[2 * o + 1] = r.r2; term[o] = r.term; agent[o] = e[l].agent;
                    log[o] = r.acted ? action : 255;
                }
            }
            warp_emit(e, m0, m1, obs + ((int64_t)t * n + first) * 117, mask + ((int64_t)t * n + first) * 54, nvalid);
        }
        for (int l = 0; l < nvalid; ++l) {
            ulonglong2 v = env_pack(e[l]);
            state[2 * (first + l)] = v.x; state[2 * (first + l) + 1] = v.y;
            if (fast) {
                st[l].steps = live_steps[l];
                st[l].sumlen = plies_start[l] + live_steps[l] - e[l].plies;
                st[l].p2w = st[l].episodes - st[l].p1w;
            }
            uint32_t a[8] = {st[l].episodes, st[l].p1w, st[l].p2w, st[l].steps, st[l].sumlen, st[l].illegal, st[l].both, st[l].maxlen};
            for (int i = 0; i < 7; ++i) stats[i] += a[i];
            stats[7] = std::max<int64_t>(stats[7], a[7]);
        }
    }
}

extern "C" __attribute__((visibility("default")))
void emul_step(unsigned long long *state, int64_t n, const int64_t *actions, uint32_t flags, int8_t *obs,
               int8_t *mask, int8_t *rew, uint8_t *term, uint8_t *trunc, uint8_t *agent, int8_t *fobs,
               int8_t *fmask, int64_t *stats) {
    const bool same_step = (flags & GBL_AUTORESET_MASK) == GBL_AUTORESET_SAME_STEP;
    for (int64_t first = 0; first < n; first += 32) {
        int nvalid = (int)std::min<int64_t>(32, n - first);
        Env e[32]; uint32_t m0[32], m1[32]; Stats st[32]; StepResult r[32];
        for (int l = 0; l < 32; ++l) {
            env_clear(e[l]); memset(&st[l], 0, sizeof(Stats));
            uint32_t action = 255;
            if (l < nvalid) {
                env_unpack(e[l], {state[2 * (first + l)], state[2 * (first + l) + 1]});
                long long a = actions[first + l];
                action = (a < 0 || a > 255) ? 254u : (uint32_t)a;
            }
            uint32_t u, up; occupancy(e[l], u, up); legal_mask(e[l].xo, e[l].yo, u, up, m0[l], m1[l]);
            r[l] = env_step<false>(e[l], m0[l], m1[l], action, flags, st[l]);
            occupancy(e[l], u, up); legal_mask(e[l].xo, e[l].yo, u, up, m0[l], m1[l]);
        }
        if (same_step) {
            if (fobs) warp_emit(e, m0, m1, fobs + first * 117, fmask + first * 54, nvalid);
            for (int l = 0; l < 32; ++l)
                if (r[l].term && !r[l].skipped) {
                    env_clear(e[l]);
                    uint32_t u, up; occupancy(e[l], u, up); legal_mask(e[l].xo, e[l].yo, u, up, m0[l], m1[l]);
                }
        }
        warp_emit(e, m0, m1, obs + first * 117, mask + first * 54, nvalid);
        for (int l = 0; l < nvalid; ++l) {
            int64_t g = first + l;
            ulonglong2 v = env_pack(e[l]);
            state[2 * g] = v.x; state[2 * g + 1] = v.y;
            rew[2 * g] = r[l].r1; rew[2 * g + 1] = r[l].r2; term[g] = r[l].term; trunc[g] = r[l].trunc; agent[g] = e[l].agent;
            uint32_t a[8] = {st[l].episodes, st[l].p1w, st[l].p2w, st[l].steps, st[l].sumlen, st[l].illegal, st[l].both, st[l].maxlen};
            for (int i = 0; i < 7; ++i) stats[i] += a[i];
            stats[7] = std::max<int64_t>(stats[7], a[7]);
        }
    }
}
