/*
 * gobblet_b200.h -- C ABI of libgobblet_b200.so (sm_100a).
 *
 * The reference (elliottower/gobblet-rl) is pure Python and has no FFI; the boundary below is the
 * set of entry points a binding for its per-step hot path would need.  Each entry cites the
 * reference interface it replaces (paths under the reference repo).  Plain pointers and sizes only:
 * no torch / C++ types cross this boundary.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the caller's current CUDA device unless it says "host";
 *   - the caller owns all buffers; the library allocates nothing;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*; NULL =
 *     legacy default stream); functions are re-entrant and keep no global mutable state;
 *   - return value 0 = success, negative = error (GBL_E_*), message via gbl_last_error()
 *     (thread-local); no C++ exception crosses the ABI;
 *   - actions outside [0,54) are treated as illegal moves, never as undefined behaviour.
 *
 * Layouts
 *   state  : n * 16 bytes, 16-byte aligned.  Two little-endian u64 per env:
 *              w0 = X1 | Y1<<27 | meta_lo<<54      w1 = X2 | Y2<<27 | meta_hi<<54
 *            Xp/Yp = 27-bit boards of player p's odd/even pieces (pieces 1,3,5 / 2,4,6 of
 *            board.py:33), bit 9*(size-1)+pos -- the reference's own `squares` index.
 *            meta (20 bits) = agent_selection | done<<1 | truncated<<2 | plies<<3 (saturating).
 *   obs    : int8 [n][3][3][13], byte pos*13+c   (gobblet.py:188-208), 16-byte aligned base
 *   mask   : int8 [n][54]                        (gobblet.py:209-213, :223-228), 16-byte aligned base
 *   rew2   : int8 [n][2] = env.rewards[player_1], env.rewards[player_2]   (gobblet.py:255-260)
 *   stats  : int64[8] accumulated with atomics: episodes, player_1 wins, player_2 wins, live steps,
 *            sum of episode lengths, illegal moves, both-line endings, max episode
 *            length (slot 7 is a max, not a sum).
 */
#ifndef GOBBLET_B200_H
#define GOBBLET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GBL_API __attribute__((visibility("default")))
#else
#define GBL_API
#endif

#define GBL_ABI_VERSION 3
#define GBL_OBS_BYTES 117
#define GBL_MASK_BYTES 54
#define GBL_STATE_BYTES 16
#define GBL_NUM_ACTIONS 54

/* flags */
#define GBL_ILLEGAL_TERMINATE 0x0u /* env(): TerminateIllegalWrapper(illegal_reward=-1), gobblet.py:114 */
#define GBL_ILLEGAL_PASS 0x1u      /* raw_env: board untouched, turn passes, board.py:125-126 + gobblet.py:244-270 */
#define GBL_AUTORESET_OFF (0u << 1)
#define GBL_AUTORESET_SAME_STEP (1u << 1)
#define GBL_AUTORESET_NEXT_STEP (2u << 1)
#define GBL_AUTORESET_MASK (3u << 1)
#define GBL_ACTION_SKIP_255 0x10u     /* gbl_step: action EXACTLY 255 leaves that env untouched (Tianshou steps a SUBSET of env ids);
                                         every other value outside [0,54) is an illegal move */
#define GBL_SLOT_FROM_ZERO 0x20u      /* gbl_rollout_random: step t goes to ring slot t % ring (not (step_base+t) % ring):
                                         the launch writes straight into the slots of a caller's trajectory buffer */
#define GBL_EMIT_INITIAL 0x40u        /* gbl_rollout_random, with GBL_SLOT_FROM_ZERO: slot 0 of obs_out / mask_out / agent_out receives the
                                         observation of the state as loaded, step t goes to slot t+1 (ring >= T+1); rew_out /
                                         term_out / final_*_out keep T slots indexed by t -- the [T+1] / [T] layout of a
                                         trajectory buffer, filled without a separate observe or copy launch */
#define GBL_NO_BULK_STORE_HINT 0x8000u /* gbl_rollout_random, 32 / 64-thread blocks only: copy the observation image with the lanes
                                         (LDS.128 -> STG.128) instead of the copy engine (tuning aid) */
#define GBL_SPLIT_HINT 0x10000u       /* gbl_rollout_random, 32-thread blocks only: two warps per 32 envs, one emits the observations,
                                       * the other the masks (both run the game logic); default for small batches (tuning aid) */
#define GBL_NO_SPLIT_HINT 0x20000u    /* never split */
#define GBL_BLOCK_HINT_SHIFT 12       /* gbl_rollout_random: bits 12-14 = threads per block, 0 auto, 1..4 = 32/64/128/256 (tuning aid) */
#define GBL_MEASURE_SKIP_OBS_STORES 0x100u  /* gbl_rollout_random only: measurement aid, do everything but the obs stores */
#define GBL_MEASURE_SKIP_MASK_STORES 0x200u /* gbl_rollout_random only: measurement aid, do everything but the mask stores */
#define GBL_STORE_DEFAULT_POLICY 0x8u /* use plain st.global instead of streaming (evict-first) stores */

/* errors */
#define GBL_E_INVALID (-1) /* bad argument (null pointer, misalignment, n < 0 ...) */
#define GBL_E_CUDA (-2)    /* CUDA runtime error; text in gbl_last_error() */

GBL_API int gbl_abi_version(void);
GBL_API const char *gbl_last_error(void);

/* raw_env.reset(): fresh Board, player_1 to move, turn 0  (gobblet.py:275-290, board.py:33) */
GBL_API int gbl_reset(void *state, int64_t n, void *stream);
/* reset the envs with which[i] != 0 (Tianshou venv.reset(ids), collector_manual_policy.py:132-145) */
GBL_API int gbl_reset_masked(void *state, const uint8_t *which, int64_t n, void *stream);

/* raw_env.observe(agent_selection) for every env: planes + live 54-way mask
 * (gobblet.py:179-215, :223-228 -> 54 x Board.is_legal, board.py:82-115). agent_id nullable. */
GBL_API int gbl_observe(const void *state, int8_t *obs, int8_t *mask, uint8_t *agent_id, int64_t n,
                void *stream);

/* env.step(action); env.last() for every env (gobblet.py:231-273; Board.play_turn board.py:118-132;
 * check_for_winner board.py:183-194; wrapper gobblet.py:110-117), then observe the new
 * agent_selection.  actions: n integers of action_bytes (1 = uint8, 4 = int32, 8 = int64) each.
 * Nullable: rew2, terminated, truncated, agent_id, final_obs/final_mask (terminal observation that
 * same-step auto-reset would otherwise replace), stats. */
GBL_API int gbl_step(void *state, const void *actions, int32_t action_bytes, int8_t *obs, int8_t *mask,
             int8_t *rew2, uint8_t *terminated, uint8_t *truncated, uint8_t *agent_id,
             int8_t *final_obs, int8_t *final_mask, int64_t *stats, int64_t n, uint32_t flags,
             void *stream);

/* Fused random-legal-action rollout: the example_basic.py:50-67 / main_random.py:23-37 loop for n
 * lockstep envs and T steps in ONE launch, state register-resident.  Action of env i at absolute
 * step s = step_base + t: j = mulhi(Philox4x32-10(key = seed, ctr = (env_id_base+i, s>>2, 0))[s&3],
 * popcount(mask)), the j-th legal action in ascending order (uniform over the mask, as
 * example_basic.py:58-61).  Per step the next observation and mask are written to ring slot
 * (step_base+t) % ring of obs_out / mask_out (slot strides in bytes, multiples of 16).
 * step_base_dev (nullable, device uint64): when given it REPLACES step_base and is read by the kernel at
 * launch, so a CUDA graph holding this launch advances through the Philox stream when the caller bumps
 * the counter inside the same graph.
 * Nullable: obs_out+mask_out (simulate only), rew_out [ring][n][2], term_out [ring][n],
 * agent_out [ring][n], action_log [T][n] (255 = no action), final_obs_out+final_mask_out (same ring and
 * slot strides as obs_out / mask_out: the observation BEFORE a same-step reset replaces it = Tianshou's
 * obs_next, collector_manual_policy.py:80-106), stats.
 * An env that arrives finished under same-step auto-reset spends its first step being reset (no action,
 * action_log 255, terminated reported once more), like gbl_step. */
GBL_API int gbl_rollout_random(void *state, int64_t n, int32_t T, uint64_t seed, uint64_t env_id_base,
                       uint64_t step_base, const uint64_t *step_base_dev, int8_t *obs_out, int8_t *mask_out,
                       int64_t obs_slot_stride, int64_t mask_slot_stride, int32_t ring,
                       int8_t *rew_out, uint8_t *term_out, uint8_t *agent_out, uint8_t *action_log,
                       int8_t *final_obs_out, int8_t *final_mask_out, int64_t *stats, uint32_t flags, void *stream);

/* ---- packed wire format: the same step for HOST-side consumers -------------------------------------------
 * The observation planes and the mask are 171 bytes of 0/1 per env; across PCIe they travel as BITS.
 * One record = 6 little-endian u32 (24 bytes) per env:
 *   w0..w3  bit i (i = pos*13 + c, the byte index of gobblet.py:188-208) of the 128-bit value = obs byte i;
 *           bits 117-118 = rewards[player_1] + 1, 119-120 = rewards[player_2] + 1, 121 = terminated,
 *           122 = truncated, 123 = agent_selection, 124-127 = 0
 *   w4, w5  bit a of the 64-bit value = action_mask[a], bits 54-63 = 0          (gobblet.py:209-213)
 * gbl_step_packed = gbl_step writing `rec` [n][6] (and `final_rec`, nullable: the terminal observation under
 * same-step auto-reset) instead of the expanded tensors: 24 instead of 176 bytes per env cross PCIe.
 * gbl_observe_packed = gbl_observe in the same format (rewards 0, flags as the env carries them). */
GBL_API int gbl_step_packed(void *state, const void *actions, int32_t action_bytes, uint32_t *rec, uint32_t *final_rec,
                    int64_t *stats, int64_t n, uint32_t flags, void *stream);
GBL_API int gbl_observe_packed(const void *state, uint32_t *rec, int64_t n, void *stream);

/* HOST side of the wire format (all pointers are HOST pointers; runs on a persistent thread pool inside the
 * library -- its only global state; calls are serialised by a mutex).  Expands n records into the
 * reference-shaped arrays obs int8 [n][3][3][13], mask int8 [n][54] (gobblet.py:179-215), rew2 int8 [n][2],
 * terminated / truncated / agent_id uint8 [n] (each nullable).  nthreads <= 0: one thread per core of the
 * calling thread's affinity mask.  AVX-512BW (64 bits -> 64 bytes per instruction, non-temporal stores) when
 * the CPU has it, a table-driven path otherwise; both produce identical bytes.
 * gbl_host_unpack_chunked additionally takes `nchunks` cudaEvent_t handles: envs [chunk_end[c-1], chunk_end[c])
 * are expanded as soon as events[c] has completed (the D2H copy of that chunk), so the expansion of one chunk
 * overlaps the PCIe transfer of the next. */
GBL_API int gbl_host_unpack(const uint32_t *rec, int64_t n, int8_t *obs, int8_t *mask, int8_t *rew2,
                    uint8_t *terminated, uint8_t *truncated, uint8_t *agent_id, int32_t nthreads);
GBL_API int gbl_host_unpack_chunked(const uint32_t *rec, int64_t n, int32_t nchunks, const int64_t *chunk_end,
                            void *const *events, int8_t *obs, int8_t *mask, int8_t *rew2,
                            uint8_t *terminated, uint8_t *truncated, uint8_t *agent_id, int32_t nthreads);
/* The whole end-to-end step for a HOST-side driver of n envs in ONE call -- `env.step(a); env.last()` per env
 * (gobblet.py:231-273, :179-215) with host actions in and host arrays out.  On `stream`: actions_host -> d_actions
 * (H2D), one gbl_step_packed launch into d_rec, then d_rec -> h_rec (D2H) in nchunks pieces (envs
 * [chunk_end[c-1], chunk_end[c]); multiples of 1024 keep the expander on its fast path) with events[c] behind
 * each piece; the host thread pool expands every piece as soon as its event has completed, so the expansion
 * overlaps the remaining PCIe traffic.
 * Host pointers: actions_host (uint8, pinned for an asynchronous copy), h_rec [n][6] (pinned), obs / mask / rew2 /
 * terminated / truncated / agent_id (any host memory; obs == mask == NULL: no expansion, the caller consumes h_rec).
 * Device pointers: state, d_actions [n], d_rec [n][6], stats.  stream / events: a cudaStream_t and nchunks
 * cudaEvent_t handles owned by the caller; the state must not be in use by other streams.  Returns after all
 * results are in host memory. */
GBL_API int gbl_step_host(void *state, const uint8_t *actions_host, int64_t n, uint32_t flags, uint8_t *d_actions,
                  uint32_t *d_rec, uint32_t *h_rec, int32_t nchunks, const int64_t *chunk_end, void *stream,
                  void *const *events, int8_t *obs, int8_t *mask, int8_t *rew2, uint8_t *terminated,
                  uint8_t *truncated, uint8_t *agent_id, int64_t *stats, int32_t nthreads);

/* measurement aids: fill `bytes` of host memory with the pool (the write bandwidth the expander is bound by),
 * the pool size a call with `nthreads` would use, and the expander variant (1 = AVX-512BW, 0 = table).
 * mode: 0 = non-temporal stores, 1 = regular stores. */
GBL_API int gbl_host_fill(void *dst, int64_t bytes, int32_t nthreads, int32_t mode);
GBL_API int gbl_host_threads(int32_t nthreads);
/* wall-clock marks (seconds since entry) of this thread's last gbl_step_host call: [0] everything enqueued,
 * [1 .. nchunks] chunk c handed to the expander, [nchunks + 1] expansion finished; returns how many there are */
GBL_API int gbl_host_last_timing(double *out, int32_t cap);
GBL_API int gbl_host_simd(void);
/* GBL_HOST_STORE_MODE: gbl_host_set_store_mode(0) = staged + non-temporal (default), 1 = direct regular stores */
GBL_API int gbl_host_set_store_mode(int32_t mode);

/* Uniform sample over each mask row with the same Philox stream as the rollout
 * (random_admissible_policy_rllib.py:23-30, example_basic.py:58-61).  act[i] = -1 for an empty row.
 * step_dev (nullable, device uint64) replaces `step` when given (see gbl_rollout_random). */
GBL_API int gbl_sample_legal(const int8_t *mask, uint64_t seed, uint64_t env_id_base, uint64_t step,
                     const uint64_t *step_dev, int32_t *act, int64_t n, void *stream);

/* GreedyGobbletPolicy(depth).compute_action for n boards, one warp per board
 * (greedy_policy.py:38-221; depth 1, 2 or 3 -- the reference's depth-3 branch, :160-208, only re-assigns the choice
 * depth 2 has already made, edits a local list and leaves its own loop, so depth 3 runs the depth-2 search:
 * tests/golden/greedy_depth3.npz records that the reference returns the same).  prev3 (nullable): int16 [n][3], the agent's last three
 * actions, -1 = none (greedy_policy.py:211-214).  Outputs (all nullable except act):
 *   act[i]     final action (Philox pick among the candidates when the fallback fires; -1 if the mask is empty)
 *   chosen[i]  choice before the random fallback, -1 = None
 *   cand[i]    bit a set <=> a in actions_depth1 (what np.random.choice draws from, :217)
 *   used_fallback[i] */
GBL_API int gbl_greedy(const int8_t *obs, const int8_t *mask, const int16_t *prev3, int32_t depth,
               uint64_t seed, uint64_t ctr_base, int32_t *act, int32_t *chosen, uint64_t *cand,
               uint8_t *used_fallback, int64_t n, void *stream);

/* Debug / interchange views of the reference's own state array (board.py:33):
 * squares int8 [n][27] signed piece numbers, agent uint8 [n] (0 = player_1 to move). */
GBL_API int gbl_export_squares(const void *state, int8_t *squares, uint8_t *agent, int64_t n, void *stream);
/* gbl_import_squares validates what Board.is_legal would reject (board.py:94-95: a piece placed twice) or
 * could never hold (a piece on a level that is not its size's): such an env is loaded as the EMPTY board
 * and, when invalid_count (nullable, device int32) is given, counted there. */
GBL_API int gbl_import_squares(void *state, const int8_t *squares, const uint8_t *agent, int64_t n,
                       int32_t *invalid_count, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GOBBLET_B200_H */
