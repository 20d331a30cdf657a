// gobblet_core.cuh -- register-resident bitboard engine for 3x3 Gobblet Gobblers (sm_100a).
//
// One environment per thread.  The per-env state is kept MOVER-RELATIVE in four registers:
//   xo, yo : 27-bit boards of the mover's odd / even pieces (pieces 1,3,5 / 2,4,6), bit 9*level+pos
//            -- the reference's `squares` index (gobblet_rl/game/board.py:33, :76-79)
//   xp, yp : same for the opponent
// so handing over the turn is a register swap.  Everything below is integer bit arithmetic; nothing is
// a contraction, so no tensor cores.
#pragma once
#include <stdint.h>

namespace gbl {

constexpr uint32_t F0 = 0x000001FFu, F1 = 0x0003FE00u, F2 = 0x07FC0000u, B27 = 0x07FFFFFFu;

struct Env {
    uint32_t xo, yo, xp, yp;  // mover-relative boards
    uint32_t agent;           // agent_selection: 0 = player_1 (gobblet.py:160-161)
    uint32_t plies;           // raw_env.turn (gobblet.py:270)
    uint32_t done, trunc;     // terminations / truncations (gobblet.py:263; wrapper :114)
};

struct Stats {
    uint32_t episodes, p1w, p2w, steps, sumlen, illegal, both, maxlen;
};

__device__ __forceinline__ void env_clear(Env &e) {  // raw_env.reset, gobblet.py:275-290
    e.xo = e.yo = e.xp = e.yp = 0;
    e.agent = 0; e.plies = 0; e.done = 0; e.trunc = 0;
}

// index of the most significant set bit, 0xFFFFFFFF for 0 (one FLO instruction)
__device__ __forceinline__ uint32_t bfind(uint32_t v) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
#else  // host-side emulation used only by tests/emul (never by the product)
    return v ? 31u - (uint32_t)__builtin_clz(v) : 0xFFFFFFFFu;
#endif
}

// ---- HBM state (16 B / env, absolute player_1 / player_2 layout, see include/gobblet_b200.h) ----
__device__ __forceinline__ void env_unpack(Env &e, ulonglong2 v) {
    uint32_t x1 = (uint32_t)v.x & B27, y1 = (uint32_t)(v.x >> 27) & B27;
    uint32_t x2 = (uint32_t)v.y & B27, y2 = (uint32_t)(v.y >> 27) & B27;
    uint32_t meta = (uint32_t)(v.x >> 54) | ((uint32_t)(v.y >> 54) << 10);
    e.agent = meta & 1u; e.done = (meta >> 1) & 1u; e.trunc = (meta >> 2) & 1u; e.plies = meta >> 3;
    if (e.agent == 0) { e.xo = x1; e.yo = y1; e.xp = x2; e.yp = y2; }
    else              { e.xo = x2; e.yo = y2; e.xp = x1; e.yp = y1; }
}

__device__ __forceinline__ ulonglong2 env_pack(const Env &e) {
    uint32_t x1, y1, x2, y2;
    if (e.agent == 0) { x1 = e.xo; y1 = e.yo; x2 = e.xp; y2 = e.yp; }
    else              { x1 = e.xp; y1 = e.yp; x2 = e.xo; y2 = e.yo; }
    uint32_t plies = e.plies > 0x1FFFFu ? 0x1FFFFu : e.plies;
    uint32_t meta = e.agent | (e.done << 1) | (e.trunc << 2) | (plies << 3);
    ulonglong2 v;
    v.x = (uint64_t)x1 | ((uint64_t)y1 << 27) | ((uint64_t)(meta & 0x3FFu) << 54);
    v.y = (uint64_t)x2 | ((uint64_t)y2 << 27) | ((uint64_t)(meta >> 10) << 54);
    return v;
}

// ---- rules -------------------------------------------------------------------------------------
// Occupancy summary shared by legality and the winner test.
//   u  field s = squares holding a piece of size >= s      (gobble rule, board.py:106-115)
//   up field s = squares holding a piece of size  > s      (check_covered, board.py:203-220)
__device__ __forceinline__ void occupancy(const Env &e, uint32_t &u, uint32_t &up) {
    uint32_t occ = e.xo | e.yo | e.xp | e.yp;
    u = occ | (occ >> 9) | (occ >> 18);
    up = u >> 9;
}

// zero the 9-bit fields of `free` whose piece (one-hot field of c) is covered
__device__ __forceinline__ uint32_t movable_fields(uint32_t covered_onehot, uint32_t free) {
    uint32_t keep = ((covered_onehot & F0) ? 0u : F0) | ((covered_onehot & F1) ? 0u : F1) |
                    ((covered_onehot & F2) ? 0u : F2);
    return free & keep;
}

// 54-way legal-action mask of the player owning (x, y): Board.is_legal for a = 0..53
// (board.py:82-115; gobblet.py:223-228).  bit a of (m1:m0), a = 9*(piece-1)+pos.
__device__ __forceinline__ void legal_mask(uint32_t x, uint32_t y, uint32_t u, uint32_t up,
                                           uint32_t &m0, uint32_t &m1) {
    uint32_t free = ~u & B27;                    // field s: top piece absent or smaller than s
    uint32_t mx = movable_fields(x & up, free);  // a covered piece may not move (board.py:101)
    uint32_t my = movable_fields(y & up, free);
    m0 = (mx & F0) | ((my & F0) << 9) | ((mx & F1) << 9) | ((my & F1) << 18);
    m1 = ((my & F1) >> 14) | ((mx & F2) >> 14) | ((my & F2) >> 5);
}

// squares whose visible (top) piece belongs to the owner of (x, y): get_flatboard, board.py:159-177
__device__ __forceinline__ uint32_t tops(uint32_t x, uint32_t y, uint32_t up) {
    uint32_t v = (x | y) & ~up;
    return (v | (v >> 9) | (v >> 18)) & 0x1FFu;
}

// Complete lines of both players at once.  Returns lm: bits 0..7 = lines of `to` complete, bits
// 16..23 = lines of `tp`, bit i = line i in the reference's order (board.py:135-153):
// (0,1,2) (3,4,5) (6,7,8) (0,3,6) (1,4,7) (2,5,8) (0,4,8) (2,4,6).
__device__ __forceinline__ uint32_t line_bits(uint32_t to, uint32_t tp) {
    uint32_t t = to | (tp << 16);
    uint32_t t1 = t >> 1, t2 = t >> 2, t3 = t >> 3, t4 = t >> 4, t6 = t >> 6, t8 = t >> 8;
    uint32_t a = t & t1 & t2 & 0x00490049u;  // bits 0,3,6 : lines 0,1,2
    uint32_t b = t & t3 & t6 & 0x00070007u;  // bits 0,1,2 : lines 3,4,5
    uint32_t d1 = t & t4 & t8 & 0x00010001u; // bit 0      : line 6
    uint32_t d2 = t & t2 & t4 & 0x00040004u; // bit 2      : line 7
    uint32_t ac = (a | (a >> 2) | (a >> 4)) & 0x00070007u;
    return ac | (b << 3) | (d1 << 6) | (d2 << 5);
}

// check_for_winner (board.py:183-194): later lines overwrite earlier ones, so the winner is the
// owner of the highest-index complete line.  A line cannot be complete for both players, hence the
// two 8-bit line sets differ in their top bit and an unsigned compare decides.
// Returns +1 (owner of `to`), -1 (owner of `tp`), 0; both = both players own a line (quirk Q1).
__device__ __forceinline__ int winner_rel(uint32_t to, uint32_t tp, bool &both) {
    uint32_t lm = line_bits(to, tp);
    uint32_t lo = lm & 0xFFu, lp = lm >> 16;
    both = lo && lp;
    return (lo > lp) - (lp > lo);
}

// Board.play_turn for a LEGAL action of the mover (board.py:118-132)
__device__ __forceinline__ void apply_move(Env &e, uint32_t action) {
    uint32_t k = (action * 57u) >> 9;   // piece-1 = action // 9   (board.py:67-68), exact for 0..53
    uint32_t pos = action - 9u * k;     // action % 9              (board.py:63-64)
    uint32_t f9 = 9u * (k >> 1);        // level offset            (board.py:71-79)
    uint32_t field = 0x1FFu << f9, bit = 1u << (f9 + pos);
    if (k & 1u) e.yo = (e.yo & ~field) | bit;   // clear the previous location, if placed (:128-130)
    else        e.xo = (e.xo & ~field) | bit;
}

// hand the turn to the other player (gobblet.py:246, :267): the views are mover-relative
__device__ __forceinline__ void pass_turn(Env &e) {
    uint32_t t;
    t = e.xo; e.xo = e.xp; e.xp = t;
    t = e.yo; e.yo = e.yp; e.yp = t;
    e.agent ^= 1u;
}

// ---- one env.step (gobblet.py:231-273 + wrapper :110-117) -----------------------------------------
struct StepResult {
    int r1, r2;       // env.rewards[player_1], [player_2]
    bool term, trunc; // flags of THIS step
    bool acted;       // an action was consumed (false for dead envs / reset-only / skipped steps)
    bool skipped;     // GBL_ACTION_SKIP_255: the env was not in the stepped subset (never reset)
};

// (m0, m1) = legal mask of the mover BEFORE the move.  kFast drops the paths a random-legal rollout
// with same-step auto-reset can never take (dead env, illegal action, skipped env) and leaves the
// statistics that follow arithmetically from the step count (steps, sumlen, p2w, illegal) to the caller.
template <bool kFast>
__device__ __forceinline__ StepResult env_step(Env &e, uint32_t m0, uint32_t m1, uint32_t action,
                                               uint32_t flags, Stats &st) {
    StepResult r = {0, 0, false, false, true, false};
    const uint32_t autoreset = (flags >> 1) & 3u;
    if (!kFast && (flags & 0x10u) && action == 255u) {   // GBL_ACTION_SKIP_255: env not in the stepped subset
        r.acted = false; r.skipped = true;
        r.term = e.done != 0; r.trunc = e.trunc != 0;
        return r;
    }
    if (!kFast && e.done) {
        r.acted = false;
        if (autoreset == 2u) env_clear(e);          // next_step: this call only resets
        else { r.term = true; r.trunc = e.trunc != 0; }
        return r;
    }
    if (!kFast) st.steps++;
    const uint32_t mover = e.agent;
    bool legal = true;
    if (!kFast) {
        uint32_t w = action < 32u ? m0 : m1;
        legal = action < 54u && ((w >> (action & 31u)) & 1u);
        if (!legal) {
            st.illegal++;
            if (!(flags & 1u)) {                    // TerminateIllegalWrapper, gobblet.py:114
                if (mover == 0) r.r1 = -1; else r.r2 = -1;
                r.term = r.trunc = true;
                e.done = 1; e.trunc = 1;
                st.episodes++; st.sumlen += e.plies; st.maxlen = max(st.maxlen, e.plies);
                return r;
            }
        }
    }
    // play_turn is a no-op for an illegal move (board.py:125-126), but check_game_over still runs on the
    // board as it is (gobblet.py:248) -- observable on imported positions that already hold a line
    if (legal) apply_move(e, action);
    bool both = false;
    uint32_t u, up;
    occupancy(e, u, up);
    const int w = winner_rel(tops(e.xo, e.yo, up), tops(e.xp, e.yp, up), both);
    pass_turn(e);
    e.plies++;
    // gobblet.py:248-263, branch-free: the winner is the mover when w > 0, the other player when w < 0
    const bool over = w != 0;
    const bool p1_won = over && ((w > 0) == (mover == 0));
    r.term = over;
    r.r1 = over ? (p1_won ? 1 : -1) : 0;
    r.r2 = -r.r1;
    e.done = over;
    st.episodes += over; st.p1w += p1_won; st.both += both;
    st.maxlen = max(st.maxlen, over ? e.plies : 0u);
    if (!kFast) { st.p2w += over && !p1_won; st.sumlen += over ? e.plies : 0u; }
    return r;
}

// ---- Philox4x32-10 counter-based generator (Salmon et al. 2011) -------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ uint4 draw_block(uint64_t seed, uint64_t env_id, uint64_t step, uint32_t tag) {
    return philox4x32_10(make_uint4((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)(step >> 2), tag),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__device__ __forceinline__ uint32_t pick_word(uint4 b, uint32_t i) {
    return i == 0 ? b.x : i == 1 ? b.y : i == 2 ? b.z : b.w;
}

// index of the j-th (0-based) set bit of the 54-bit mask (m1:m0), j < popcount
__device__ __forceinline__ uint32_t select_bit(uint32_t m0, uint32_t m1, uint32_t j) {
    uint32_t c = __popc(m0);
    bool hi = j >= c;
    uint32_t w = hi ? m1 : m0, pos = hi ? 32u : 0u;
    j -= hi ? c : 0u;
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) {
        c = __popc(w & ((1u << sh) - 1u));
        bool up = j >= c;
        j -= up ? c : 0u;
        w = up ? w >> sh : w;
        pos += up ? sh : 0;
    }
    return pos;
}

// uniform legal action: j = floor(draw * count / 2^32).  kNonEmpty: the caller knows the mask has a set bit (every
// valid position has >= 10 legal actions), so the empty-mask guard -- a divergent region around the search -- goes.
template <bool kNonEmpty = false>
__device__ __forceinline__ uint32_t sample_action(uint32_t m0, uint32_t m1, uint32_t draw) {
    uint32_t cnt = __popc(m0) + __popc(m1);
    if (kNonEmpty) return select_bit(m0, m1, __umulhi(draw, cnt));
    return cnt ? select_bit(m0, m1, __umulhi(draw, cnt)) : 0u;
}

// ---- emission: per-env boards -> coalesced int8 tensors ----------------------------------------------
// A warp owns 32 consecutive envs => 32*117 = 3744 contiguous observation bytes and 32*54 = 1728
// contiguous mask bytes (both multiples of 16), staged per warp in shared memory:
//   * observation: the image is the FINAL byte layout (byte lane*117 + pos*13 + c, gobblet.py:188-208).
//     It is kept all-zero between steps; each lane SCATTERS the <= 12 one-bytes of its pieces (one FLO
//     + one IMAD + one predicated STS.U8 per piece) and, for player_2, the 9 bytes of plane 12; one lane
//     then hands the whole image to the copy engine (TMA bulk store) while the warp expands the mask, and
//     the warp re-zeroes the image once the engine has read it.
//   * mask: each lane shifts its 54-bit mask to its bit offset in the warp's packed stream (boundary
//     words merged with one shuffle); every lane then turns 16 stream bits into 16 bytes
//     (nibble * 0x00204081 & 0x01010101).
// Per warp and step: one 3744-byte bulk store + 108 full 128-bit, fully coalesced STG for the mask
// (234 + 108 STG with -DGBL_BULK_STORE=0 and in the ragged last warp).
// Protocol: [stage_recycle] | stage_env | __syncwarp | emit_chunk | __syncwarp, repeated; see the comments of each.
constexpr int PART_BOTH = 0, PART_OBS = 1, PART_MASK = 2;   // which output stream(s) a warp stages and emits
constexpr int OBS_IMG_BYTES = 32 * 117, MASK_WORDS = 54;
constexpr int STAGE_BYTES = OBS_IMG_BYTES + 4 * MASK_WORDS + 8;  // 3968, multiple of 16
constexpr int OBS_VEC = 234, MASK_VEC = 108;                    // uint4 stores per warp

struct LaneCfg {
    uint32_t mfo, mso;
    bool mn2;
};

__device__ __forceinline__ LaneCfg make_lane_cfg(uint32_t lane) {
    LaneCfg c;
    uint32_t m = 54u * lane, m2 = m + 54u;
    c.mfo = m >> 5; c.mso = m & 31u; c.mn2 = ((m2 >> 5) - c.mfo) == 2u;
    return c;
}

__device__ __forceinline__ uint4 expand16(uint32_t h) {  // 16 bits -> 16 bytes of 0/1
    const uint32_t M = 0x00204081u, K = 0x01010101u;
    uint4 v;
    v.x = ((h & 0xFu) * M) & K;
    v.y = (((h >> 4) & 0xFu) * M) & K;
    v.z = (((h >> 8) & 0xFu) * M) & K;
    v.w = (((h >> 12) & 0xFu) * M) & K;
    return v;
}

template <bool kStreaming>
__device__ __forceinline__ void store16(int8_t *p, uint4 v) {
    if (kStreaming) __stcs(reinterpret_cast<uint4 *>(p), v);
    else *reinterpret_cast<uint4 *>(p) = v;
}

// zero the observation image once per kernel (every emit_chunk leaves it zeroed again)
__device__ __forceinline__ void stage_init(uint8_t *stage, uint32_t lane) {
    uint4 *img = reinterpret_cast<uint4 *>(stage);
    for (uint32_t q = lane; q < OBS_VEC; q += 32u) img[q] = make_uint4(0u, 0u, 0u, 0u);
}

// store the byte 1 at shared-memory offset `off` of `base` iff cond != 0, as ONE predicated STS.U8
// (a C++ `if` around the store compiles to a divergent branch with BSSY/BSYNC per piece)
__device__ __forceinline__ void store_one_if(uint8_t *base, uint32_t off, uint32_t cond) {
#ifdef __CUDA_ARCH__
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(base) + off;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u8 [%1], %2;\n\t}"
                 :: "r"(cond), "r"(addr), "r"(1u) : "memory");
#else  // host-side emulation used only by tests/emul (never by the product)
    if (cond) base[off] = 1;
#endif
}

// store a 32-bit word to shared memory iff cond, as one predicated STS (no divergent region)
__device__ __forceinline__ void store_word_if(uint32_t *p, uint32_t v, bool cond) {
#ifdef __CUDA_ARCH__
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u32 [%1], %2;\n\t}"
                 :: "r"((uint32_t)cond), "r"(addr), "r"(v) : "memory");
#else  // host-side emulation used only by tests/emul (never by the product)
    if (cond) *p = v;
#endif
}

// one byte per piece on the board: plane c of the mover's piece c+1 at 0..5, opponent 6..11
__device__ __forceinline__ void scatter_pieces(uint8_t *mine, uint32_t w, int plane0) {
#pragma unroll
    for (int f = 0; f < 3; ++f) {
        const uint32_t fld = w & (0x1FFu << (9 * f));                                  // one-hot or empty
        store_one_if(mine, 13u * bfind(fld) + (uint32_t)(plane0 + 2 * f - 117 * f), fld);   // pos = bfind - 9f
    }
}

// TMA bulk store (cp.async.bulk, SASS UBLKCP.G.S) of the observation image: one lane hands the 3744-byte
// image to the copy engine instead of 32 lanes looping LDS.128 -> STG.128.  Compile-time choice (build.py).
#ifndef GBL_BULK_STORE
#define GBL_BULK_STORE 1
#endif
#ifndef GBL_BULK_EVICT_FIRST
#define GBL_BULK_EVICT_FIRST 0       // L2 evict-first hint on the bulk store: measured equal (4.034e10 vs 4.034e10 env-steps/s)
#endif
#if defined(__CUDA_ARCH__) && GBL_BULK_STORE
constexpr bool kBulkStore = true;
#else
constexpr bool kBulkStore = false;
#endif

__device__ __forceinline__ void bulk_store_issue(void *gdst, const void *ssrc, uint32_t bytes) {
#ifdef __CUDA_ARCH__
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ssrc);
#if GBL_BULK_EVICT_FIRST
    uint64_t pol;                        // write-once output: ask L2 to evict these lines first
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" :: "l"(gdst), "r"(saddr), "r"(bytes), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(saddr), "r"(bytes) : "memory");
#endif
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_store_wait_read() {   // the shared-memory source may be overwritten
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_store_wait_all() {    // the global writes have been performed
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_store_wait_all_but_latest() {   // every group but the newest has completed
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
#endif
}
__device__ __forceinline__ void fence_smem_for_bulk() {    // generic-proxy smem writes -> visible to the copy engine
#ifdef __CUDA_ARCH__
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// Bulk path only: take the observation image back from the copy engine (the bulk store issued by the previous
// emit_chunk must have READ it) and zero it for the next scatter.  Callers put it right before stage_env, i.e.
// AFTER the register-only game logic of the next step: the engine reads the 3744 bytes while the warp computes
// (waiting right behind the mask stores -- the first version -- cost a third of all warp stall samples).
template <bool kBulk = kBulkStore, int kPart = PART_BOTH>
__device__ __forceinline__ void stage_recycle(uint8_t *stage, uint32_t lane) {
    if (!kBulk || kPart == PART_MASK) return;
    if (lane == 0) bulk_store_wait_read();
    __syncwarp();
    uint4 *img = reinterpret_cast<uint4 *>(stage);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t q = lane + 32u * i;
        if (i < 7 || q < OBS_VEC) img[q] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
}

// stage: all 32 lanes of the warp must call (shuffle inside).  stage = this warp's STAGE_BYTES.
// The caller puts a __syncwarp() between stage_env and emit_chunk and one after emit_chunk.
// kPart: PART_BOTH, or only the observation / only the mask (two warps share the emission of 32 envs, see rollout_kernel).
template <bool kBulk = kBulkStore, int kPart = PART_BOTH>
__device__ __forceinline__ void stage_env(uint8_t *stage, const LaneCfg &c, uint32_t lane, const Env &e,
                                          uint32_t m0, uint32_t m1) {
    if (kPart != PART_MASK) {
        uint8_t *mine = stage + 117u * lane;
        scatter_pieces(mine, e.xo, 0);
        scatter_pieces(mine, e.yo, 1);
        scatter_pieces(mine, e.xp, 6);
        scatter_pieces(mine, e.yp, 7);
#pragma unroll
        for (int p = 0; p < 9; ++p) mine[13 * p + 12] = (uint8_t)e.agent;   // plane 12: the viewer is player_2 (gobblet.py:199-206);
                                                                             // unconditional (the image is zero): no divergent region
    }
    if (kPart != PART_OBS) {
        uint32_t *mbits = reinterpret_cast<uint32_t *>(stage + OBS_IMG_BYTES);
        const uint32_t v0 = m0 << c.mso, v1 = __funnelshift_l(m0, m1, c.mso), v2 = __funnelshift_l(m1, 0u, c.mso);
        uint32_t tail = c.mn2 ? v2 : v1;                    // partial word shared with the next lane
        uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, tail, 1);
        if (lane == 0) prev = 0;
        mbits[c.mfo] = v0 | prev;
        store_word_if(mbits + c.mfo + 1, v1, c.mn2);
    }
    if (kBulk && kPart != PART_MASK) fence_smem_for_bulk();
}

// copy / expand + store.  obs_chunk / mask_chunk point at the warp's first env; nvalid = envs of this
// warp that exist (32 except in the last warp).  opts: bit 0 / 1 = GBL_MEASURE_SKIP_OBS / MASK_STORES;
// EMIT_REUSE_RING / EMIT_REUSE_ALWAYS = the same global slot is written again later in this launch (ring < T):
// PTX orders separate bulk async-groups only through wait_group, so the store that last targeted this slot
// must have COMPLETED (not merely been read) before the next one is issued -- that is the group before the
// newest one when ring >= 2, the newest one when ring == 1.
constexpr uint32_t EMIT_REUSE_RING = 4u, EMIT_REUSE_ALWAYS = 8u;
template <bool kStreaming, bool kBulk = kBulkStore, int kPart = PART_BOTH>
__device__ __forceinline__ void emit_chunk(uint8_t *stage, uint32_t lane, int8_t *obs_chunk,
                                           int8_t *mask_chunk, int nvalid, uint32_t opts = 0u) {
    const uint32_t skip = opts & 3u;
    constexpr bool kObs = kPart != PART_MASK, kMask = kPart != PART_OBS;
    uint4 *img = reinterpret_cast<uint4 *>(stage);
    const uint16_t *hb = reinterpret_cast<const uint16_t *>(stage + OBS_IMG_BYTES);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    if (nvalid == 32) {
        if (kBulk && kObs) {
            if (lane == 0 && !(skip & 1u)) {
                if (opts & EMIT_REUSE_ALWAYS) bulk_store_wait_all();
                else if (opts & EMIT_REUSE_RING) bulk_store_wait_all_but_latest();
                bulk_store_issue(obs_chunk, stage, OBS_IMG_BYTES);
            }
        } else if (kObs) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t q = lane + 32u * i;
                if (i < 7 || q < OBS_VEC) {
                    uint4 v = img[q];
                    img[q] = zero;
                    store16<kStreaming>(obs_chunk + 16u * q, v);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t q = lane + 32u * i;
            if (kMask && (i < 3 || q < MASK_VEC) && !(skip & 2u)) store16<kStreaming>(mask_chunk + 16u * q, expand16(hb[q]));
        }
        // kBulk: the image stays with the copy engine; stage_recycle() takes it back before the next stage_env
    } else {  // ragged last warp: vector stores while fully inside, bytes at the edge
        const uint32_t ob = 117u * nvalid, mb = 54u * nvalid;
        for (uint32_t q = lane; kObs && q < OBS_VEC; q += 32u) {
            uint4 v = img[q];
            img[q] = zero;
            if (16u * q + 16u <= ob) store16<kStreaming>(obs_chunk + 16u * q, v);
            else
                for (uint32_t j = 0; j < 16u && 16u * q + j < ob; ++j) {
                    uint32_t w = j < 4 ? v.x : j < 8 ? v.y : j < 12 ? v.z : v.w;
                    obs_chunk[16u * q + j] = (int8_t)(w >> (8u * (j & 3u)));
                }
        }
        for (uint32_t q = lane; kMask && q < MASK_VEC; q += 32u) {
            uint4 v = expand16(hb[q]);
            if (16u * q + 16u <= mb) store16<kStreaming>(mask_chunk + 16u * q, v);
            else
                for (uint32_t j = 0; j < 16u && 16u * q + j < mb; ++j) {
                    uint32_t w = j < 4 ? v.x : j < 8 ? v.y : j < 12 ? v.z : v.w;
                    mask_chunk[16u * q + j] = (int8_t)(w >> (8u * (j & 3u)));
                }
        }
    }
}

}  // namespace gbl
