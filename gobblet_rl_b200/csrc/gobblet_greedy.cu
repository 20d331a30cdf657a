// gobblet_greedy.cu -- warp-per-board GreedyGobbletPolicy (depth 1 / 2; depth 3 == depth 2, see gbl_greedy), sm_100a.
//
// Follows gobblet_rl/game/greedy_policy.py:38-221.  The 54 candidate moves (depth 1) and, per surviving root
// move, the 54 opponent replies (depth 2) are evaluated one per lane in two rounds of 32; warp votes turn the
// per-lane verdicts into the few facts the reference's order-dependent list logic depends on.
//
// What keeps the instruction count down (the kernel is issue-bound, not memory-bound):
//  * INCREMENTAL winner test.  check_for_winner only looks at who owns the top of each square.  A move takes
//    one piece from the top of its origin square o (exposing what was under it) to the top of its target p,
//    so with the 9-bit maps  tm / to (squares whose top is mine / theirs)  and  um / uo (squares whose SECOND
//    piece from the top is mine / theirs)  of the position before the move
//        mover's tops  = (t_mover & ~o | u_mover & o) | p        other's tops = (t_other | u_other & o) & ~p
//    -- five logic ops and two 512-byte table look-ups (complete-line sets) per evaluated position instead of
//    replaying the move on the bitboards.
//  * The per-root facts that do not depend on the lane (the opponent's legal replies after the root move and
//    the four maps above) are computed lane-parallel, one root per lane, parked in shared memory and fetched
//    back with ONE broadcast LDS.128 per root.
//  * The list logic needs, for almost every board, only "can the opponent win at once after root a" (one vote)
//    and "do all replies win for me" (one vote): the LAST root that is safe decides (greedy_policy.py:153-157).
//    The block move (:141-143) can only be chosen when no root is safe and depth 1 found no win; that case
//    re-walks the unsafe roots with full 64-bit reply sets (slow path, rare).
//  * A block works through several boards per warp, so the line table is staged once per block.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/gobblet_b200.h"
#include "gobblet_core.cuh"

namespace gbl {

#ifndef GBL_GREEDY_MIN_BLOCKS
#define GBL_GREEDY_MIN_BLOCKS 5     // resident blocks per SM asked of the compiler (48 registers; 4 / 5 / 6 measured: 123 / 120 / 149 us)
#endif
constexpr int GREEDY_BLOCK = 256, GREEDY_WARPS = GREEDY_BLOCK / 32;
constexpr uint32_t FULL = 0xFFFFFFFFu;

// Complete-line sets for every 9-bit "squares I own on top" mask: bit i = line i of board.py:135-153 is
// complete.  The 8-line test (26 ALU instructions) is a 512-byte shared-memory lookup here.
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

__device__ __forceinline__ uint64_t ballot64(bool lo, bool hi) {
    return (uint64_t)__ballot_sync(FULL, lo) | ((uint64_t)__ballot_sync(FULL, hi) << 32);
}

// who owns the top (t*) and the second piece from the top (u*) of every square: get_flatboard, board.py:159-177
struct Tops { uint32_t tm, to, um, uo; };
__device__ __forceinline__ Tops summarize(uint32_t me, uint32_t th) {      // me / th: 27-bit boards (all pieces of a side)
    const uint32_t occ = me | th, o2 = (occ >> 18) & 0x1FFu, o1 = (occ >> 9) & 0x1FFu, x21 = o2 ^ o1;
    const uint32_t m2 = (me >> 18) & 0x1FFu, m1 = (me >> 9) & 0x1FFu, m0 = me & 0x1FFu;
    const uint32_t t2 = (th >> 18) & 0x1FFu, t1 = (th >> 9) & 0x1FFu, t0 = th & 0x1FFu;
    Tops s;
    s.tm = m2 | (m1 & ~o2) | (m0 & ~o2 & ~o1);
    s.to = t2 | (t1 & ~o2) | (t0 & ~o2 & ~o1);
    s.um = (o2 & m1) | (m0 & x21);      // large over medium -> the medium one; exactly one of large / medium over small
    s.uo = (o2 & t1) | (t0 & x21);
    return s;
}

// check_for_winner (board.py:183-194) after the MOVER took a piece from origin bit o (0 = not yet placed) to
// target bit p: +1 mover / -1 other / 0.  The owner of the highest-index complete line wins = the larger set.
__device__ __forceinline__ void tops_after(uint32_t t_mover, uint32_t u_mover, uint32_t t_other, uint32_t u_other,
                                           uint32_t o, uint32_t p, uint32_t &mover, uint32_t &other) {
    mover = ((t_mover & ~o) | (u_mover & o)) | p;
    other = (t_other | (u_other & o)) & ~p;
}

__global__ void __launch_bounds__(GREEDY_BLOCK, GBL_GREEDY_MIN_BLOCKS)
greedy_kernel(const int8_t *__restrict__ obs, const int8_t *__restrict__ mask, const int16_t *__restrict__ prev3,
              int32_t depth, uint64_t seed, uint64_t ctr_base, int32_t *act, int32_t *chosen_out,
              uint64_t *cand_out, uint8_t *fallback_out, int64_t n, int32_t boards_per_warp) {
    // the table is 512-byte aligned: a 9-bit index is OR-ed into its shared-memory address, and that OR folds into
    // the LOP3 that produces the index (no address add per look-up)
    __shared__ __align__(512) uint8_t lut[512];
    __shared__ __align__(16) uint32_t rootinfo[GREEDY_WARPS][56][8];
    reinterpret_cast<uint16_t *>(lut)[threadIdx.x] = reinterpret_cast<const uint16_t *>(kLineLut.v)[threadIdx.x];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t (*const info)[8] = rootinfo[warp];
    uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(lut);
    asm volatile("mov.u32 %0, %0;" : "+r"(lut_s));            // opaque: keeps the compiler from re-deriving the address in the root loop
    auto lines = [lut_s](uint32_t tops9) -> uint32_t {        // complete-line set of a 9-bit top map: one LDS.U8
        uint32_t v;
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(lut_s | tops9));
        return v;
    };

    // ---- per-lane constants: the actions this lane evaluates in round 0 / 1 (a = lane, lane + 32) ----------
    uint32_t pbit[2], fsh[2], is_y[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint32_t a = lane + 32u * r, k = (a * 57u) >> 9, pos = a - 9u * k;     // board.py:63-68
        pbit[r] = a < 54u ? 1u << pos : 0u;
        fsh[r] = 9u * (k >> 1);                                                     // level of the piece (:71-79)
        is_y[r] = k & 1u;
    }
    // bit pos*13+c of the observation -> packed board bit t = 27*w + 9*level + pos (w: my odd / my even / their odd / their even)
    uint32_t src_word[4], src_bit[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t t = 32u * r + lane, w = t / 27u, i = t - 27u * w, f = i / 9u, p = i - 9u * f;
        const uint32_t bit = 13u * p + (w & 1u) + 2u * f + 6u * (w >> 1);
        src_word[r] = t < 108u ? bit >> 5 : 4u;
        src_bit[r] = bit & 31u;
    }

    const int64_t first = ((int64_t)blockIdx.x * GREEDY_WARPS + warp) * boards_per_warp;
    for (int64_t b = first; b < first + boards_per_warp && b < n; ++b) {
        // ---- observation planes -> bitmap (bit pos*13+c), coalesced byte loads + ballots -------------
        const int8_t *o = obs + b * GBL_OBS_BYTES;
        uint32_t s[5];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t i = 32u * r + lane;                               // clamped index: no divergent branch around the load
            s[r] = __ballot_sync(FULL, (r < 3 || i < GBL_OBS_BYTES) && o[r < 3 ? i : min(i, (uint32_t)GBL_OBS_BYTES - 1u)] != 0);
        }
        s[4] = 0;
        const int8_t *mk = mask + b * GBL_MASK_BYTES;
        const uint64_t maskbits = ballot64(mk[lane] != 0, lane + 32u < GBL_MASK_BYTES && mk[min(lane + 32u, (uint32_t)GBL_MASK_BYTES - 1u)] != 0);

        // ---- bitmap -> boards of "me" (planes 0-5) and "them" (planes 6-11)   (greedy_policy.py:43-71)
        uint32_t pk[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t sw = src_word[r];
            const uint32_t word = sw == 0 ? s[0] : sw == 1 ? s[1] : sw == 2 ? s[2] : sw == 3 ? s[3] : 0u;
            pk[r] = __ballot_sync(FULL, (word >> src_bit[r]) & 1u);
        }
        const uint32_t xo = pk[0] & B27;
        const uint32_t yo = __funnelshift_r(pk[0], pk[1], 27) & B27;          // bits 27..53
        const uint32_t xp = __funnelshift_r(pk[1], pk[2], 22) & B27;          // bits 54..80
        const uint32_t yp = __funnelshift_r(pk[2], pk[3], 17) & B27;          // bits 81..107

        uint32_t l0, l1;
        {
            const uint32_t occ = xo | yo | xp | yp, u = occ | (occ >> 9) | (occ >> 18);
            legal_mask(xo, yo, u, u >> 9, l0, l1);
        }
        const uint64_t legal0 = (uint64_t)l0 | ((uint64_t)l1 << 32);         // my legal moves on the ORIGINAL board
        const Tops t0 = summarize(xo | yo, xp | yp);

        // ---- depth 1 (greedy_policy.py:84-101): one candidate per lane, two rounds ------------------
        const uint64_t keys = maskbits & legal0;                              // `results` keys, if reached
        uint32_t org_me[2], org_op[2], tm1[2], to1[2];
        bool win[2], loss[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            org_me[r] = ((is_y[r] ? yo : xo) >> fsh[r]) & 0x1FFu;            // where my piece of this action stands (one-hot / 0)
            org_op[r] = ((is_y[r] ? yp : xp) >> fsh[r]) & 0x1FFu;            // same for the opponent's piece (replies)
            tops_after(t0.tm, t0.um, t0.to, t0.uo, org_me[r], pbit[r], tm1[r], to1[r]);
            const bool cand_here = (keys >> (lane + 32u * r)) & 1ull;
            const uint32_t lm = lines(tm1[r]), lt = lines(to1[r]);
            win[r] = cand_here && lm > lt;
            loss[r] = cand_here && lt > lm;
        }
        const uint64_t win1 = ballot64(win[0], win[1]), loss1 = ballot64(loss[0], loss[1]);

        uint64_t cand = maskbits;                                             // actions_depth1 (:77-79)
        int ncand = __popcll(cand), chosen = -1, stop = 64;
        for (uint64_t ev = win1 | loss1; ev; ev &= ev - 1) {
            const int a = __ffsll((long long)ev) - 1;
            if ((win1 >> a) & 1ull) { chosen = a; stop = a; break; }         // :92-94
            if (ncand > 1) { cand &= ~(1ull << a); --ncand; }                // :95-99
            else { stop = a; break; }                                        // :100-101
        }
        const uint64_t reached = stop == 64 ? ~0ull : ((2ull << stop) - 1ull);
        const uint64_t roots = keys & reached & ~win1 & ~loss1;              // results[a] == 0

        // ---- depth 2 (greedy_policy.py:103-157) ----------------------------------------------------------
        if (depth > 1 && roots) {
            // lane-parallel, one root per lane and round: the opponent's legal replies after the root move
            // (greedy_policy.py:110-114) and the top / under maps of that position
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint32_t a = lane + 32u * r;
                if ((roots >> a) & 1ull) {
                    const uint32_t field = 0x1FFu << fsh[r], bit = pbit[r] << fsh[r];
                    const uint32_t x1 = is_y[r] ? xo : (xo & ~field) | bit, y1 = is_y[r] ? (yo & ~field) | bit : yo;
                    const uint32_t occ = x1 | y1 | xp | yp, u1 = occ | (occ >> 9) | (occ >> 18);
                    uint32_t rl, rh;
                    legal_mask(xp, yp, u1, u1 >> 9, rl, rh);
                    const Tops t1 = summarize(x1 | y1, xp | yp);
                    *reinterpret_cast<uint4 *>(info[a]) = make_uint4(rl, rh, t1.tm, t1.to);     // every fact in its own word:
                    *reinterpret_cast<uint2 *>(info[a] + 4) = make_uint2(t1.um, t1.uo);         // no unpacking in the root loop
                }
            }
            __syncwarp();
            const int ncand_d1 = ncand;
            int safe = -1;
            uint64_t unsafe = 0;
            bool done = false;
#pragma unroll 1
            for (int half = 0; half < 2 && !done; ++half) {                  // 32-bit root sets: cheaper loop control than 64-bit
                for (uint32_t rs = half ? (uint32_t)(roots >> 32) : (uint32_t)roots; rs; rs &= rs - 1u) {
                    const int a = 32 * half + __ffs((int)rs) - 1;
                    const uint4 ri = *reinterpret_cast<const uint4 *>(info[a]);          // broadcast LDS.128 + LDS.64
                    const uint2 rj = *reinterpret_cast<const uint2 *>(info[a] + 4);
                    const uint32_t rtm = ri.z, rto = ri.w, rum = rj.x, ruo = rj.y;
                    const uint32_t rep[2] = {ri.x, ri.y};
                    bool opp_wins = false, not_mine = false;
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        uint32_t tt, tm;
                        tops_after(rto, ruo, rtm, rum, org_op[r], pbit[r], tt, tm);  // the opponent is the mover
                        const uint32_t lt = lines(tt), lm = lines(tm);
                        const bool is_reply = (rep[r] >> lane) & 1u;
                        opp_wins = __any_sync(FULL, is_reply && lt > lm);
                        if (opp_wins) break;                                 // unsafe root: nothing else matters here
                        not_mine |= is_reply && !(lm > lt);
                    }
                    if (opp_wins) {                                          // :131-136
                        unsafe |= 1ull << a;
                        if (ncand > 1) { cand &= ~(1ull << a); --ncand; }
                    } else {
                        safe = a;                                            // :153-157  the last safe root wins ...
                        if (!__any_sync(FULL, not_mine)) { done = true; break; }   // :146-151  ... unless every reply wins for me
                    }
                }
            }
            if (safe >= 0) chosen = safe;
            else if (chosen < 0) {
                // no safe root and no depth-1 win: the block move (:141-143) -- the opponent's first winning reply
                // that is also a legal move of mine, looked for root by root while more than one candidate is left
                int nc = ncand_d1;
                for (uint64_t rs = unsafe; rs && nc > 1; rs &= rs - 1) {
                    const int a = __ffsll((long long)rs) - 1;
                    const uint4 ri = *reinterpret_cast<const uint4 *>(info[a]);
                    const uint2 rj = *reinterpret_cast<const uint2 *>(info[a] + 4);
                    const uint32_t rtm = ri.z, rto = ri.w, rum = rj.x, ruo = rj.y;
                    const uint32_t rep[2] = {ri.x, ri.y};
                    uint32_t theirs[2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        uint32_t tt, tm;
                        tops_after(rto, ruo, rtm, rum, org_op[r], pbit[r], tt, tm);
                        theirs[r] = __ballot_sync(FULL, ((rep[r] >> lane) & 1u) && lines(tt) > lines(tm));
                    }
                    const uint64_t W = (uint64_t)theirs[0] | ((uint64_t)theirs[1] << 32);
                    --nc;                                                    // this root left the candidate list
                    const uint64_t elig = nc > 1 ? W : (W & (0 - W));        // later replies only while len > 1
                    const uint64_t blk = elig & legal0;
                    if (blk) { chosen = __ffsll((long long)blk) - 1; break; }
                }
            }
            __syncwarp();                                                    // info[] is rewritten by the next board
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
}

}  // namespace gbl

extern "C" int gbl__set_error(const char *msg);  // gobblet_engine.cu (thread-local message of gbl_last_error)

extern "C" int gbl_greedy(const int8_t *obs, const int8_t *mask, const int16_t *prev3, int32_t depth, uint64_t seed,
                          uint64_t ctr_base, int32_t *act, int32_t *chosen, uint64_t *cand, uint8_t *used_fallback,
                          int64_t n, void *stream) {
    if (n < 0 || depth < 1 || depth > 3) { gbl__set_error("gbl_greedy: n < 0 or depth not in {1,2,3}"); return GBL_E_INVALID; }
    if (depth == 3) depth = 2;      // greedy_policy.py:160-208 cannot change the result of depth 2 (include/gobblet_b200.h)
    if (n == 0) return 0;
    if (!obs || !mask || !act) { gbl__set_error("gbl_greedy: obs/mask/act must be non-null"); return GBL_E_INVALID; }
    // a few boards per warp once the grid fills the machine (the block stages the line table once per 8 * bpw boards),
    // but still several waves of blocks so that the hardware block scheduler evens out the uneven boards
    // (65 536 boards: 1 / 2 / 3 / 4 / 6 / 8 boards per warp measured 121 / 117 / 117 / 118 / 119 / 125 us)
    const int64_t warps_4_waves = 148 * GBL_GREEDY_MIN_BLOCKS * gbl::GREEDY_WARPS * 4;
    int32_t bpw = (int32_t)(n / warps_4_waves);
    bpw = bpw < 1 ? 1 : bpw > 4 ? 4 : bpw;
    if (const char *env = getenv("GBL_GREEDY_BPW")) bpw = atoi(env) > 0 ? atoi(env) : bpw;       // tuning knob
    const int64_t per_block = (int64_t)gbl::GREEDY_WARPS * bpw;
    const unsigned grid = (unsigned)((n + per_block - 1) / per_block);
    gbl::greedy_kernel<<<grid, gbl::GREEDY_BLOCK, 0, (cudaStream_t)stream>>>(obs, mask, prev3, depth, seed, ctr_base, act,
                                                                             chosen, cand, used_fallback, n, bpw);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { gbl__set_error(cudaGetErrorString(e)); return GBL_E_CUDA; }
    return 0;
}
