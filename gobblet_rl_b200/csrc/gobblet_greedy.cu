// gobblet_greedy.cu -- warp-per-board GreedyGobbletPolicy (depth 1 / 2), sm_100a.
//
// Follows gobblet_rl/game/greedy_policy.py:38-221.  The 54 candidate moves (depth 1) and, per
// surviving root move, the 54 opponent replies (depth 2) are evaluated one per lane in two rounds
// of 32; __ballot_sync turns the per-lane verdicts into 64-bit sets on which the reference's
// order-dependent list logic (candidate pruning, "block" move, last-safe-move-wins) is resolved
// with a handful of scalar bit operations executed uniformly by the whole warp.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gobblet_b200.h"
#include "gobblet_core.cuh"

namespace gbl {

constexpr int GREEDY_BLOCK = 256, GREEDY_WARPS = GREEDY_BLOCK / 32;
constexpr uint32_t FULL = 0xFFFFFFFFu;

__device__ __forceinline__ void apply_xy(uint32_t &x, uint32_t &y, uint32_t action) {  // board.py:118-132
    uint32_t k = (action * 57u) >> 9, pos = action - 9u * k, f9 = 9u * (k >> 1);
    uint32_t field = 0x1FFu << f9, bit = 1u << (f9 + pos);
    if (k & 1u) y = (y & ~field) | bit; else x = (x & ~field) | bit;
}

// Complete-line sets for every 9-bit "squares I own on top" mask: bit i = line i of board.py:135-153 is
// complete.  The search evaluates ~3000 positions per board and the integer pipe is its bottleneck (ncu:
// ALU 80 % busy, LSU idle), so the 8-line test (26 ALU instructions) is a 512-byte shared-memory lookup here.
struct LineLut { uint8_t v[512]; };
constexpr LineLut make_line_lut() {
    LineLut t{};
    const int L[8][3] = {{0, 1, 2}, {3, 4, 5}, {6, 7, 8}, {0, 3, 6}, {1, 4, 7}, {2, 5, 8}, {0, 4, 8}, {2, 4, 6}};
    for (int m = 0; m < 512; ++m) {
        int b = 0;
        for (int l = 0; l < 8; ++l)
            if (((m >> L[l][0]) & (m >> L[l][1]) & (m >> L[l][2])) & 1) b |= 1 << l;
        t.v[m] = (uint8_t)b;
    }
    return t;
}
__device__ const LineLut kLineLut = make_line_lut();

// check_for_winner after a move, relative to (mine, theirs): +1 / -1 / 0   (board.py:183-194): the owner
// of the highest-index complete line wins, i.e. the larger of the two line sets (they share no line)
__device__ __forceinline__ int winner_of(const uint8_t *lut, uint32_t xm, uint32_t ym, uint32_t xt, uint32_t yt) {
    uint32_t occ = xm | ym | xt | yt, u = occ | (occ >> 9) | (occ >> 18), up = u >> 9;
    uint32_t lm = lut[tops(xm, ym, up)], lt = lut[tops(xt, yt, up)];
    return (lm > lt) - (lt > lm);
}

__device__ __forceinline__ uint64_t ballot64(bool lo, bool hi) {
    return (uint64_t)__ballot_sync(FULL, lo) | ((uint64_t)__ballot_sync(FULL, hi) << 32);
}

__global__ void __launch_bounds__(GREEDY_BLOCK)
greedy_kernel(const int8_t *__restrict__ obs, const int8_t *__restrict__ mask, const int16_t *__restrict__ prev3,
              int32_t depth, uint64_t seed, uint64_t ctr_base, int32_t *act, int32_t *chosen_out,
              uint64_t *cand_out, uint8_t *fallback_out, int64_t n) {
    __shared__ __align__(16) uint8_t lut[512];
    reinterpret_cast<uint16_t *>(lut)[threadIdx.x] = reinterpret_cast<const uint16_t *>(kLineLut.v)[threadIdx.x];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * GREEDY_WARPS + (threadIdx.x >> 5);
    if (b >= n) return;

    // ---- observation planes -> bitmap (bit pos*13+c), coalesced byte loads + ballots -------------
    const int8_t *o = obs + b * GBL_OBS_BYTES;
    uint32_t s[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t i = 32u * r + lane;
        s[r] = __ballot_sync(FULL, i < GBL_OBS_BYTES && o[i] != 0);
    }
    const int8_t *mk = mask + b * GBL_MASK_BYTES;
    const uint64_t maskbits = ballot64(mk[lane] != 0, lane + 32u < GBL_MASK_BYTES && mk[lane + 32u] != 0);

    // ---- bitmap -> boards of "me" (planes 0-5) and "them" (planes 6-11)   (greedy_policy.py:43-71)
    // packed index t = 27*w + 9*level + pos, w: 0 = my odd pieces, 1 = my even, 2 = their odd, 3 = their even
    uint32_t pk[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t t = 32u * r + lane, w = t / 27u, i = t - 27u * w, f = i / 9u, p = i - 9u * f;
        uint32_t bit = 13u * p + (w & 1u) + 2u * f + 6u * (w >> 1);
        uint32_t word = bit < 32u ? s[0] : bit < 64u ? s[1] : bit < 96u ? s[2] : s[3];
        pk[r] = __ballot_sync(FULL, t < 108u && ((word >> (bit & 31u)) & 1u));
    }
    const uint32_t xo = pk[0] & B27;
    const uint32_t yo = __funnelshift_r(pk[0], pk[1], 27) & B27;          // bits 27..53
    const uint32_t xp = __funnelshift_r(pk[1], pk[2], 22) & B27;          // bits 54..80
    const uint32_t yp = __funnelshift_r(pk[2], pk[3], 17) & B27;          // bits 81..107

    uint32_t u, up, l0, l1;
    {
        uint32_t occ = xo | yo | xp | yp;
        u = occ | (occ >> 9) | (occ >> 18);
        up = u >> 9;
    }
    legal_mask(xo, yo, u, up, l0, l1);
    const uint64_t legal0 = (uint64_t)l0 | ((uint64_t)l1 << 32);         // my legal moves on the ORIGINAL board

    // ---- depth 1 (greedy_policy.py:84-101): one candidate per lane, two rounds ------------------
    bool win[2], loss[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        uint32_t a = lane + 32u * r;
        int w = 0;
        if (a < 54u && ((legal0 & maskbits) >> a) & 1ull) {
            uint32_t x = xo, y = yo;
            apply_xy(x, y, a);
            w = winner_of(lut, x, y, xp, yp);
        }
        win[r] = w > 0; loss[r] = w < 0;
    }
    const uint64_t win1 = ballot64(win[0], win[1]), loss1 = ballot64(loss[0], loss[1]);
    const uint64_t keys = maskbits & legal0;                              // `results` keys, if reached

    uint64_t cand = maskbits;                                             // actions_depth1 (:77-79)
    int ncand = __popcll(cand), chosen = -1, stop = 64;
    for (uint64_t ev = keys & (win1 | loss1); ev; ev &= ev - 1) {
        int a = __ffsll((long long)ev) - 1;
        if ((win1 >> a) & 1ull) { chosen = a; stop = a; break; }         // :92-94
        if (ncand > 1) { cand &= ~(1ull << a); --ncand; }                // :95-99
        else { stop = a; break; }                                        // :100-101
    }
    const uint64_t reached = stop == 64 ? ~0ull : ((2ull << stop) - 1ull);
    uint64_t roots = keys & reached & ~win1 & ~loss1;                    // results[a] == 0

    // ---- depth 2 (greedy_policy.py:103-157): loop roots, replies in lanes -------------------------
    if (depth > 1) {
        // The opponent's legal replies after each root move (greedy_policy.py:110-114) do not depend on the
        // lane, so computing them inside the root loop would repeat ~45 instructions on all 32 lanes per root.
        // Each lane computes them for "its" two roots instead (lane, lane+32) and the loop broadcasts them.
        uint32_t rep_lo[2], rep_hi[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t a = lane + 32u * r;
            rep_lo[r] = rep_hi[r] = 0;
            if ((roots >> a) & 1ull) {
                uint32_t x1 = xo, y1 = yo;
                apply_xy(x1, y1, a);
                uint32_t occ = x1 | y1 | xp | yp, u1 = occ | (occ >> 9) | (occ >> 18);
                legal_mask(xp, yp, u1, u1 >> 9, rep_lo[r], rep_hi[r]);
            }
        }
        for (; roots; roots &= roots - 1) {
            const int a = __ffsll((long long)roots) - 1;
            uint32_t x1 = xo, y1 = yo;
            apply_xy(x1, y1, (uint32_t)a);
            const uint32_t r0 = __shfl_sync(FULL, a < 32 ? rep_lo[0] : rep_lo[1], a & 31);
            const uint32_t r1 = __shfl_sync(FULL, a < 32 ? rep_hi[0] : rep_hi[1], a & 31);
            const uint64_t replies = (uint64_t)r0 | ((uint64_t)r1 << 32);
            uint32_t theirs[2] = {0u, 0u}, notmine[2] = {0u, 0u};     // ballots of the two rounds of 32 replies
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t a2 = lane + 32u * r;
                int w = 1;
                const bool is_reply = (replies >> a2) & 1ull;
                if (is_reply) {
                    uint32_t x2 = xp, y2 = yp;
                    apply_xy(x2, y2, a2);
                    w = winner_of(lut, x1, y1, x2, y2);
                }
                theirs[r] = __ballot_sync(FULL, is_reply && w < 0);
                notmine[r] = __ballot_sync(FULL, is_reply && w <= 0);
                // once a move has been chosen only "can the opponent win at all" matters for this root (the block
                // move of :141-143 needs `chosen is None`), so a hit in the first round settles it
                if (r == 0 && chosen >= 0 && theirs[0]) break;
            }
            const uint64_t W = (uint64_t)theirs[0] | ((uint64_t)theirs[1] << 32);   // replies that win for the opponent
            const bool all_mine = (notmine[0] | notmine[1]) == 0;                    // vacuously true without replies
            if (W) {
                if (ncand > 1) {                                          // :131-136
                    if ((cand >> a) & 1ull) { cand &= ~(1ull << a); --ncand; }
                    const uint64_t elig = ncand > 1 ? W : (W & (0 - W));  // later replies only while len > 1
                    const uint64_t blk = elig & legal0;                   // :141-143
                    if (chosen < 0 && blk) chosen = __ffsll((long long)blk) - 1;
                }
            } else {
                chosen = a;                                               // :146-157
                if (all_mine) break;
            }
        }
    }

    // ---- repetition rule + random fallback (greedy_policy.py:211-219) -----------------------------
    bool fb = chosen < 0;
    if (!fb && prev3) {
        const int16_t *p = prev3 + 3 * b;
        fb = chosen == p[0] || chosen == p[1] || chosen == p[2];
    }
    int final_act = chosen;
    if (fb) {
        if (ncand > 0) {
            uint4 d = draw_block(seed, ctr_base + (uint64_t)b, 0ull, 1u);
            final_act = (int)select_bit((uint32_t)cand, (uint32_t)(cand >> 32), __umulhi(d.x, (uint32_t)ncand));
        } else final_act = -1;
    }
    if (lane == 0) {
        act[b] = final_act;
        if (chosen_out) chosen_out[b] = chosen;
        if (cand_out) cand_out[b] = cand;
        if (fallback_out) fallback_out[b] = fb;
    }
}

}  // namespace gbl

extern "C" int gbl__set_error(const char *msg);  // gobblet_engine.cu (thread-local message of gbl_last_error)

extern "C" int gbl_greedy(const int8_t *obs, const int8_t *mask, const int16_t *prev3, int32_t depth, uint64_t seed,
                          uint64_t ctr_base, int32_t *act, int32_t *chosen, uint64_t *cand, uint8_t *used_fallback,
                          int64_t n, void *stream) {
    if (n < 0 || depth < 1 || depth > 2) { gbl__set_error("gbl_greedy: n < 0 or depth not in {1,2}"); return GBL_E_INVALID; }
    if (n == 0) return 0;
    if (!obs || !mask || !act) { gbl__set_error("gbl_greedy: obs/mask/act must be non-null"); return GBL_E_INVALID; }
    const unsigned grid = (unsigned)((n + gbl::GREEDY_WARPS - 1) / gbl::GREEDY_WARPS);
    gbl::greedy_kernel<<<grid, gbl::GREEDY_BLOCK, 0, (cudaStream_t)stream>>>(obs, mask, prev3, depth, seed, ctr_base, act,
                                                                             chosen, cand, used_fallback, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { gbl__set_error(cudaGetErrorString(e)); return GBL_E_CUDA; }
    return 0;
}
