/*
 * envforge_b200 -- C ABI of the B200-native batched-stepping engine for the data-parallel hot path of
 * mariusdgm/rl-envs-forge (GridWorld, KArmedBandit, CartPole, PendulumDisk).
 *
 * The reference has no FFI layer: its boundary for this path is the gymnasium Env protocol implemented by four
 * Python classes.  Each entry point below replaces the per-instance Python method cited next to it with one call
 * that advances N independent instances held as structure-of-arrays device buffers.
 *
 * Conventions
 *   - Plain C types only.  Every buffer is DEVICE memory (pinned host memory for the h_ arguments of the *_step_host
 *     entry points) owned by the caller (PyTorch in the shipped host code);
 *     the library never allocates, frees or retains device memory and keeps no global state (one exception, stated
 *     where it is declared: the 1 KB peer mailbox of ef_peer_comm).  Descriptor structs
 *     (ef_grid_desc, ef_pendulum_params, ef_bandit_desc) are small HOST structs passed by pointer and copied into
 *     kernel parameters; pointers inside them are device pointers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are asynchronous and
 *     stream-ordered; buffers must stay alive until the stream has passed the launch.  Re-entrant; one device per
 *     call (the caller's current device).
 *   - Return value: EF_OK (0) or an ef_status code; never throws, never exits.  `ef_status_string` names a code.
 *   - Global env index g = env_offset + i (must stay < 2^32).  All in-kernel randomness is Philox4x32-10 with
 *     key = seed and counter = (g, 64-bit counter, stream id | block << 8); see csrc/ef_common.cuh (twin: oracle/philox.py).  Results therefore
 *     do not depend on how envs are sharded over launches, ranks or GPUs.
 *   - `stats` (nullable) points at EF_STATS_REPLICAS x EF_STATS_STRIDE doubles that the kernels ADD to (atomically).
 *     Replica r lives at stats + r * EF_STATS_STRIDE; a CTA adds into replica (blockIdx.x % EF_STATS_REPLICAS) so that
 *     CTAs finishing together do not serialise on one L2 atomic unit.  The statistics are the SUM over replicas of
 *       [0] finished episodes  [1] sum of their lengths  [2] sum of their returns  [3] sum of squared returns
 *       [4] env-steps executed [5..] reserved.  Integer-valued entries stay exact below 2^53.
 *     The (summed) vector is what gets all-reduced (NCCL, sum) across GPUs.
 *   - `err` (nullable) points at EF_ERR_LEN uint32, 8-byte aligned: [0] number of env-steps the reference would have
 *     raised on (atomic add); [1] reserved; [2..3] one 64-bit word describing the offender with the largest global index
 *     (atomic max): (1 + global env index) << 32 | ef_step_error kind << 24 | detail (24 bits: table row cell*4+action,
 *     or the action value).  Offending envs are left untouched (the reference raises before mutating state).
 *   - `clock` (nullable) points at 2 uint64 in device memory.  clock[0] counts steps in units of 2^-EF_CLOCK_SHIFT:
 *     (clock[0] >> EF_CLOCK_SHIFT) is ADDED to the `t` / `t0` argument to form the Philox time coordinate, and a launch
 *     that advances the envs by k steps (1 per step call, K per rollout) adds k << EF_CLOCK_SHIFT in total, every CTA its
 *     share as it retires; clock[1] is reserved.  It exists so a captured CUDA graph can be replayed without the host
 *     re-baking `t` into the kernel parameters.  Initialise it to (steps already taken) << EF_CLOCK_SHIFT.
 */
#ifndef ENVFORGE_B200_H
#define ENVFORGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EF_ABI_VERSION 2
#define EF_STATS_LEN 8
#define EF_STATS_REPLICAS 32
#define EF_STATS_STRIDE 16
#define EF_ERR_LEN 4
#define EF_CLOCK_SHIFT 20

typedef enum ef_status {
    EF_OK = 0,
    EF_ERR_INVALID_ARGUMENT = 1,   /* null/unaligned pointer, bad size, bad descriptor */
    EF_ERR_CUDA = 2,               /* a CUDA runtime call failed; see ef_last_cuda_error() */
    EF_ERR_UNSUPPORTED = 3         /* descriptor outside what the kernels support (e.g. > 65536 cells) */
} ef_status;

typedef enum ef_step_error {
    EF_STEP_OK = 0,
    EF_STEP_BAD_ACTION = 1,        /* GridWorld: Action(a) ValueError (grid_world.py:505); bandit: assert (k_armed_bandit.py:186);
                                      CartPole / PendulumDisk: assert action_space.contains(action) (cart_pole.py:108,
                                      pendulum_disk.py:109) -- |a| > 3 or NaN; detail = bits 8-31 of the fp32 action */
    EF_STEP_NO_OUTCOMES = 2,       /* GridWorld: "No outcomes defined" ValueError (grid_world.py:508-511), e.g. agent on a wall */
    EF_STEP_BAD_PROBS = 3          /* GridWorld: numpy choice "probabilities do not sum to 1" (grid_world.py:515) */
} ef_step_error;

int ef_abi_version(void);
const char* ef_status_string(int status);
/* cudaError_t of the most recent failed CUDA call on this host thread (0 if none) and its name. */
int ef_last_cuda_error(void);
const char* ef_last_cuda_error_string(void);
/* SM count / compute capability of the current device. */
int ef_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ------------------------------------------------------------------------------------------------------------
 * GridWorld   (reference: rl_envs_forge/envs/grid_world/grid_world.py)
 *
 * The host compiles the reference's dict MDP (build_mdp :245-274) into a flat device table of ef_grid_outcome,
 * indexed [(cell*4 + action) * n_slots + outcome_index], cell = row*cols + col.  Rows with a single outcome
 * replicate it over all slots, so one set of thresholds serves the whole table:
 *     outcome_index = min((r >= thr[0]) + (r >= thr[1]) + (r >= thr[2]), idx_cap)        (n_slots == 4)
 * which equals numpy's searchsorted(cumsum(p)/cumsum(p)[-1], r * 2^-32, 'right') used by Generator.choice
 * (grid_world.py:514-515); thr[k] = ceil(cdf_k * 2^32) clipped to 2^32-1, idx_cap = #thresholds below 2^32.
 * Rows the reference cannot execute carry EF_GRID_NO_OUTCOMES / EF_GRID_BAD_PROBS in every slot and MUST be compiled
 * inert: next = the row's own cell, reward = 0, EF_GRID_DONE clear (the kernels apply them and mask the step counter).
 * ------------------------------------------------------------------------------------------------------------ */
#define EF_GRID_DONE 0x1u        /* outcome lands on a terminal state */
#define EF_GRID_NO_OUTCOMES 0x2u /* row has no MDP entry (reference raises ValueError) */
#define EF_GRID_BAD_PROBS 0x4u   /* row's probabilities fail numpy's sum-to-1 check (reference raises ValueError) */

typedef struct ef_grid_outcome {
    uint16_t next;   /* destination cell */
    uint16_t flags;  /* EF_GRID_* */
    float reward;
} ef_grid_outcome;

typedef struct ef_grid_desc {
    int32_t rows, cols;
    int32_t n_slots;             /* 1 (every row deterministic) or 4 */
    int32_t idx_cap;             /* 0..3 */
    uint32_t thr[3];
    int32_t start_cell;
    int32_t episode_limit;       /* <= 0: no limit (episode_length_limit=None) */
    int32_t table_cells;         /* cells the table covers: 0 or rows*cols, or more when the host appends pseudo-cells
                                    (an off-grid start_state the reference accepts in reset, grid_world.py:541-548) */
    const ef_grid_outcome* table; /* device, table_cells*4*n_slots entries, 8-byte aligned */
} ef_grid_desc;

/* GridWorld.step (:495-528) + GridWorld.reset (:530-551) fused as same-step auto-reset, for n envs.
 *   pos/ep_len/ep_ret  in/out state: cell index, episode_length_counter, running episode return (fp32).
 *                      PACKED LAYOUT: ep_len may be NULL in every GridWorld entry point; `pos` then holds one 32-bit word
 *                      per env, cell in bits 0-15 and episode_length_counter in bits 16-31 (8 fewer bytes moved per
 *                      env-step).  Accepted only where the counter cannot pass 65535: autoreset != 0 and
 *                      0 < episode_limit <= 65535; EF_ERR_INVALID_ARGUMENT otherwise.
 *   action             [n] int32 in {0,1,2,3}, or NULL: uniform random policy from the Philox policy word.
 *   draw               [n] uint32 injected outcome draws r (u = r*2^-32), or NULL: Philox transition word of step t.
 *   obs                [n,2] int32 (row, col) AFTER auto-reset (nullable).  final_obs: (row, col) BEFORE it (nullable).
 *   reward/terminated/truncated  [n] outputs (each nullable).
 *   autoreset          0: reproduce the reference step-for-step (no reset; terminal self-loops, sticky truncation).
 */
int ef_gridworld_step(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                      const int32_t* action, const uint32_t* draw, uint64_t seed, uint64_t t, uint64_t env_offset,
                      int32_t* obs, float* reward, uint8_t* terminated, uint8_t* truncated, int32_t* final_obs,
                      double* stats, uint32_t* err, uint64_t* clock, int64_t n, int32_t autoreset, void* stream);

/* The same call with its arguments in one struct: a host loop that steps the same buffers over and over (the usual RL
 * loop) fills the struct once and per step only updates `action`, `t` and `stream` -- for a Python caller that is one
 * ctypes argument to convert instead of twenty (about 4 us of a ~10 us call). */
typedef struct ef_gridworld_step_args {
    const ef_grid_desc* grid;
    int32_t* pos;
    int32_t* ep_len;
    float* ep_ret;
    const int32_t* action;
    const uint32_t* draw;
    uint64_t seed, t, env_offset;
    int32_t* obs;
    float* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    int32_t* final_obs;
    double* stats;
    uint32_t* err;
    uint64_t* clock;
    int64_t n;
    int32_t autoreset;
    int32_t wire;     /* flags.  EF_WIRE_U8: action / obs / final_obs point at uint8 data (ef_gridworld_step_u8), else int32;
                         EF_STEP_COLD_INPUTS: a hint, see below */
    void* stream;
} ef_gridworld_step_args;
#define EF_WIRE_U8 0x1
/* Hint for the vector step kernel, results are the same either way: the env set's arrays and clock word have most likely
 * been pushed out of L2 since the set's last step (several env sets stepped in rotation, together larger than L2).  The
 * kernel then warms the clock word with a line prefetch instead of a load ahead of griddepcontrol.wait (a load would hold
 * its warp at the wait until DRAM answers): measured 6.98 vs 7.25 us per 2^20-env launch with eight sets in rotation, and
 * the other way round (6.85 vs 6.74 us) when one L2-resident set is stepped repeatedly. */
#define EF_STEP_COLD_INPUTS 0x2
int ef_gridworld_step_v(const ef_gridworld_step_args* args);

/* ef_gridworld_step in the NARROW WIRE FORMAT of a host-side actor loop: actions are uint8 [n] (values > 3 are rejected
 * like any other invalid Action), observations and final_obs are uint8 [n,2] -- needs rows <= 256 and cols <= 256
 * (EF_ERR_UNSUPPORTED otherwise).  The reference's step() takes one Python int and returns a (row, col) tuple
 * (grid_world.py:495-528); int32 is this engine's canonical widening of those, uint8 the narrowest dtype that still holds
 * them.  An env-step then moves 9 bytes between host and device (1 in; 2 + 4 + 1 + 1 out) instead of 18, which is what a
 * step() on HOST buffers is bound by (PCIe), and 25 instead of 34 through HBM.  State, randomness, statistics, error
 * reporting and every other argument are those of ef_gridworld_step: the two calls produce the same trajectories.
 * Vector path: action 4-byte aligned, obs / final_obs 8-byte aligned (else one env per thread). */
int ef_gridworld_step_u8(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret, const uint8_t* action,
                         const uint32_t* draw, uint64_t seed, uint64_t t, uint64_t env_offset, uint8_t* obs,
                         float* reward, uint8_t* terminated, uint8_t* truncated, uint8_t* final_obs, double* stats,
                         uint32_t* err, uint64_t* clock, int64_t n, int32_t autoreset, void* stream);

/* ef_gridworld_step for a batch that mixes several grid LAYOUTS (walls / terminal states / special transitions / start
 * state differ per layout; rows, cols, n_slots, thresholds and episode_limit are shared): `grid->table` holds n_layouts
 * consecutive tables of table_cells*4*n_slots entries each, env i walks table layout_index[i] (or
 * (env_offset + i) % n_layouts when layout_index is NULL) and auto-resets to start_cells[layout] (grid->start_cell is
 * ignored).  All layouts are staged in shared memory when they fit (96 KB), else looked up through L1/L2.
 * One env per thread; every other argument as in ef_gridworld_step. */
int ef_gridworld_step_layouts(const ef_grid_desc* grid, int32_t n_layouts, const int32_t* layout_index,
                              const int32_t* start_cells, int32_t* pos, int32_t* ep_len, float* ep_ret,
                              const int32_t* action, const uint32_t* draw, uint64_t seed, uint64_t t, uint64_t env_offset,
                              int32_t* obs, float* reward, uint8_t* terminated, uint8_t* truncated, int32_t* final_obs,
                              double* stats, uint32_t* err, uint64_t* clock, int64_t n, int32_t autoreset, void* stream);

/* K fused steps of a mixed-layout batch (ef_gridworld_rollout / ef_gridworld_rollout_traj for the batches of
 * ef_gridworld_step_layouts): state in registers for all K steps, auto-reset to the env's own start cell, table offset
 * and start cell fetched once per launch.  obs [K,n,2] int32 / rewards [K,n] / flags [K,n] each nullable. */
int ef_gridworld_rollout_layouts(const ef_grid_desc* grid, int32_t n_layouts, const int32_t* layout_index,
                                 const int32_t* start_cells, int32_t* pos, int32_t* ep_len, float* ep_ret,
                                 const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                 int32_t* obs, float* rewards, uint8_t* flags, double* stats, uint32_t* err,
                                 uint64_t* clock, int64_t n, void* stream);

/* GridWorld.reset for a mixed-layout batch: env i returns to start_cells[layout of env i]; mask nullable. */
int ef_gridworld_reset_layouts(int32_t n_layouts, const int32_t* layout_index, const int32_t* start_cells, int32_t cols,
                               int32_t* pos, int32_t* ep_len, float* ep_ret, int32_t* obs, const uint8_t* mask,
                               uint64_t env_offset, int64_t n, void* stream);

/* Stream-exact variant of ef_gridworld_step: the outcome draw of env i comes from ITS OWN numpy-compatible generator,
 * Generator(PCG64(SeedSequence(seed_i))) as built by gym.utils.seeding.np_random (grid_world.py:69) and consumed by
 * Generator.choice (:515), so a batch seeded with the reference's seeds reproduces the reference's trajectories without
 * injected draws.
 *   rng_state / rng_inc  [n,2] uint64 (high word, low word) PCG64 state / increment, 16-byte aligned; state is updated.
 *   thr53 / idx_cap53    host pointer to 3 thresholds T_k = ceil(cdf_k * 2^53) and the index cap: with
 *                        m = next_uint64 >> 11, outcome_index = min(sum(m >= T_k), idx_cap53)  (ignored for n_slots == 1,
 *                        where a draw is still consumed, as the reference does).
 * Rejected env-steps consume no draw (the reference raises before drawing). */
int ef_gridworld_step_pcg64(const ef_grid_desc* grid, const uint64_t* thr53, int32_t idx_cap53, int32_t* pos,
                            int32_t* ep_len, float* ep_ret, uint64_t* rng_state, const uint64_t* rng_inc,
                            const int32_t* action, uint64_t env_offset, int32_t* obs, float* reward, uint8_t* terminated,
                            uint8_t* truncated, int32_t* final_obs, double* stats, uint32_t* err, int64_t n,
                            int32_t autoreset, void* stream);

/* K fused steps (t0 .. t0+K-1) with state held in registers and auto-reset always on.
 *   actions  [K,n] int32 or NULL (random policy).   rewards [K,n] float / flags [K,n] uint8 (bit0 terminated,
 *   bit1 truncated) or NULL: nothing is written per step and only state + stats leave the kernel. */
int ef_gridworld_rollout(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                         const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                         float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock, int64_t n,
                         void* stream);

/* ef_gridworld_rollout that also materialises the trajectory of observations: obs [K,n,2] int32 (row, col) after each
 * step's auto-reset (what K step() calls would have returned); rewards / flags as above, each nullable. */
int ef_gridworld_rollout_traj(const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                              const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                              int32_t* obs, float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock,
                              int64_t n, void* stream);

/* GridWorld.reset (:530-551) for all envs (mask == NULL) or those with mask[i] != 0. */
int ef_gridworld_reset(int32_t start_cell, int32_t cols, int32_t* pos, int32_t* ep_len, float* ep_ret,
                       int32_t* obs, const uint8_t* mask, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * CartPole / PendulumDisk   (reference: rl_envs_forge/envs/inverted_pendulum/{cart_pole/cart_pole.py,
 * pendulum_disk/pendulum_disk.py}).  fp32 on device; the host derives the constants in float64.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ef_pendulum_params {
    float tau;                   /* cart_pole.py:39 / pendulum_disk.py:38 */
    /* CartPole: c[0]=gravity c[1]=pole_mass_length c[2]=1/total_mass c[3]=length c[4]=4/3 c[5]=mass_pole/total_mass
     *           c[6]=pole_mass_length/total_mass c[7]=x_threshold (cart_pole.py:56-64, 117-123)
     * PendulumDisk: c[0]=m*g*l/J c[1]=(b+K^2/R)/J c[2]=K/(R*J) c[3]=15*pi (pendulum_disk.py:58-64, 115-120) */
    float c[8];
    float angle_threshold;       /* 0.2 (theta_threshold_radians) | pi */
    float angle_termination;     /* < 0: None */
    int32_t reward_mode;         /* 0 discrete, 1 continuous linear, 2 continuous nonlinear */
    float reward_lo, reward_hi;  /* reward_angle_range */
    float reward_decay;          /* reward_decay_rate */
    float reward_inv_span;       /* 1/angle_threshold (linear) */
    float reward_inv_norm;       /* 1/(1-exp(-decay*angle_threshold)) (nonlinear) */
    int32_t episode_limit;       /* episode_length_limit */
    int32_t fixed_init;          /* 1: reset to init_state; 0: U(init_lo, init_hi) per component (_initialize_state) */
    float init_lo, init_hi;
    float init_state[4];
} ef_pendulum_params;

/* CartPole.step (cart_pole.py:107-180) + reset (:91-105) as same-step auto-reset.
 *   state   [n,4] fp32 (x, x_dot, theta, theta_dot), 16-byte aligned.   action [n] fp32 force, or NULL: U(-3,3) policy.
 *   obs     [n,4] post-reset observation (nullable; may equal `state`).  final_obs [n,4] pre-reset (nullable).
 *   acc     [n,2] (x_acc, theta_acc) info values (nullable).
 *   err     (nullable) the reference asserts action_space.contains(action) (:108): an env-step whose action is NaN or
 *           outside [-3, 3] is counted there (EF_STEP_BAD_ACTION) and leaves the env untouched -- reward 0, flags 0. */
int ef_cartpole_step(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret, const float* action,
                     uint64_t seed, uint64_t t, uint64_t env_offset, float* obs, float* reward, uint8_t* terminated,
                     uint8_t* truncated, float* final_obs, float* acc, double* stats, uint32_t* err, uint64_t* clock, int64_t n,
                     int32_t autoreset, void* stream);
int ef_cartpole_rollout(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                        const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                        float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock, int64_t n, void* stream);
/* ..._rollout that also materialises the observation trajectory: obs [K,n,4] fp32, the state after each step's auto-reset. */
int ef_cartpole_rollout_traj(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                             const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K, float* obs,
                             float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock, int64_t n, void* stream);
/* reset: initial state from params (fixed) or Philox STREAM_RESET_USER at counter `epoch`; mask nullable. */
int ef_cartpole_reset(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret, uint64_t seed,
                      uint64_t epoch, uint64_t env_offset, const uint8_t* mask, int64_t n, void* stream);

/* PendulumDisk.step (pendulum_disk.py:108-172) + reset (:94-106).  state [n,2] fp32 (alpha, alpha_dot), 8-byte
 * aligned; acc [n] alpha_dot_dot (nullable). */
int ef_pendulum_disk_step(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                          const float* action, uint64_t seed, uint64_t t, uint64_t env_offset, float* obs,
                          float* reward, uint8_t* terminated, uint8_t* truncated, float* final_obs, float* acc,
                          double* stats, uint32_t* err, uint64_t* clock, int64_t n, int32_t autoreset, void* stream);
int ef_pendulum_disk_rollout(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                             const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                             float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock, int64_t n, void* stream);
int ef_pendulum_disk_rollout_traj(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                                  const float* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K,
                                  float* obs /* [K,n,2] */, float* rewards, uint8_t* flags, double* stats, uint32_t* err, uint64_t* clock,
                                  int64_t n, void* stream);
int ef_pendulum_disk_reset(const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret, uint64_t seed,
                           uint64_t epoch, uint64_t env_offset, const uint8_t* mask, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * KArmedBandit   (reference: rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py)
 * Arms are shared by all n bandits; a default arm (family EF_ARM_DEFAULT_NORMAL) has a per-env mean re-derived
 * in-kernel from Philox STREAM_BANDIT_MEANS (the reference draws it with randn in _create_default_arms :160-165).
 * Time-varying parameters (Arm.update_param :53-64) are evaluated by the host into `arms[t - t0]`.
 * ------------------------------------------------------------------------------------------------------------ */
typedef enum ef_arm_family {
    EF_ARM_NORMAL = 0,      /* p0 mean, p1 std      (Arm.sample :76-79)   */
    EF_ARM_UNIFORM = 1,     /* p0 low,  p1 high     (:81-84)              */
    EF_ARM_EXPONENTIAL = 2, /* p0 scale             (:86-88)              */
    EF_ARM_GAMMA = 3,       /* p0 shape, p1 scale   (:90-97)              */
    EF_ARM_BETA = 4,        /* p0 a, p1 b           (:99-106)             */
    EF_ARM_LOGISTIC = 5,    /* p0 loc, p1 scale     (:108-111)            */
    EF_ARM_WEIBULL = 6,     /* p0 a                 (:113-117)            */
    EF_ARM_LOGNORMAL = 7,   /* p0 mean, p1 sigma    (:119-122)            */
    EF_ARM_DEFAULT_NORMAL = 8 /* mean = per-env default mean + p0, std p1 (_create_default_arms :160-165) */
} ef_arm_family;

typedef struct ef_arm {
    int32_t family;
    float p0, p1;
    int32_t reserved;
} ef_arm;

/* Arm parameters in float64, for the stream-exact mode below (the float64 legacy samplers take them unrounded). */
typedef struct ef_arm64 {
    int32_t family;   /* ef_arm_family */
    int32_t reserved;
    double p0, p1;
} ef_arm64;

/* KArmedBandit.step (:175-200) for n bandits over K consecutive timesteps (K = 1 is the plain step).
 *   arms     device [K,k] ef_arm: parameters in force when the reward of timestep index t0+j is drawn.
 *   action   [K,n] int32 or NULL (uniform random arm from the Philox policy word).
 *   obs      [K,n] int32 (= action, :200) nullable.  reward [K,n] fp32.  (done == 1, truncated == 0 always.)
 *   t0       number of steps taken since the last reset() -- the counter of the per-(env, step) primitive stream.
 *            (The reference re-creates RandomState(seed) in reset (:215), so its reward stream replays after a
 *            reset while `timestep`, which only drives the parameter shift, keeps counting; the host mirrors that
 *            by restarting t0 at 0 on reset and evaluating `arms` at the true timestep.) */
int ef_bandit_step(const ef_arm* arms, int32_t k, const int32_t* action, uint64_t seed,
                   uint64_t t0, uint64_t env_offset, int32_t K, int32_t* obs, float* reward, double* stats,
                   uint32_t* err, int64_t n, void* stream);
/* The same step with the default arms' per-env means LOOKED UP instead of re-derived: `means` is the fp32 [n,k] table
 * ef_bandit_default_means fills for the same (k, seed, env_offset, n) -- row i belongs to env i of THIS call -- or NULL
 * (then this is ef_bandit_step).  Bit-identical rewards either way; with two default arms in ten the table saves a
 * Philox call + Box-Muller on a fifth of the steps (40 B of HBM per env for k = 10).  Reference: the means are drawn
 * once per instance in _create_default_arms (k_armed_bandit.py:160-165) and kept in the Arm objects. */
int ef_bandit_step_means(const ef_arm* arms, int32_t k, const int32_t* action, const float* means, uint64_t seed,
                         uint64_t t0, uint64_t env_offset, int32_t K, int32_t* obs, float* reward, double* stats,
                         uint32_t* err, int64_t n, void* stream);
/* STREAM-EXACT KArmedBandit.step: every bandit owns the generator the reference builds for it -- numpy's legacy
 * RandomState(seed_i), MT19937 (k_armed_bandit.py:145, :215) -- and Arm.sample (:66-122) runs numpy's legacy float64
 * algorithms on it, so the reward stream of env i is that of KArmedBandit(seed=seed_i) without injected draws.
 *   arms       device [k] ef_arm64, parameters in force for this step (EF_ARM_DEFAULT_NORMAL: mean = means[i,arm] + p0).
 *   key        [624, n] uint32, word-major MT19937 state;  pos / has_gauss [n] int32;  gauss [n] double -- the fields of
 *              RandomState.get_state(), updated in place.  The host seeds them with numpy itself.
 *   means      [n, k] double: randn(k) of each env's generator (_create_default_arms :160-165), nullable.
 *   reward     [n] fp32 (the engine's public reward dtype), reward64 [n] double (the unrounded sample); each nullable.
 * 2.5 KB of state per env: meant for small n; the Philox kernels above are the throughput path. */
int ef_bandit_step_mt19937(const ef_arm64* arms, int32_t k, const int32_t* action, uint32_t* key, int32_t* pos,
                           int32_t* has_gauss, double* gauss, const double* means, uint64_t env_offset, int32_t* obs,
                           float* reward, double* reward64, double* stats, uint32_t* err, int64_t n, void* stream);

/* Per-env default-arm means (fp32 [n,k], the values the kernel re-derives); for KArmedBandit.arms / state. */
int ef_bandit_default_means(int32_t k, uint64_t seed, uint64_t env_offset, float* means, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * ACML envs   (reference: rl_envs_forge/envs/acml/{gambler_problem/gambler_problem.py, car_rental/car_rental.py})
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ef_gambler_params {
    int32_t goal_amount;      /* gambler_problem.py:24 */
    int32_t start_capital;    /* :26, the state reset() restores (:75) */
    int32_t episode_limit;    /* <= 0: none (the reference has none; > 0 adds truncation for auto-reset users) */
    int32_t reserved;
    uint64_t win_threshold;   /* ceil(win_probability * 2^32) in [0, 2^32]: `np.random.rand() < p` (:51) on u = r*2^-32 */
} ef_gambler_params;

/* GamblersProblem.step (:36-66) + reset (:68-76) as same-step auto-reset, for n envs.
 *   state    [n] int32 capital (in/out).   action [n] int32 bet, or NULL: uniform legal bet 0..min(s, goal-s).
 *   draw     [n] uint32 injected draws r (u = r*2^-32), or NULL: Philox transition word of step t.
 *   obs      [n] int32 capital after auto-reset; final_obs before it; reward in {0, 1}; truncated only with a limit.
 *   A bet outside [0, goal_amount] is the reference's `assert self.action_space.contains(action)` (:46): counted in
 *   `err` (EF_STEP_BAD_ACTION), env untouched.  A bet above the capital is legal in the reference and here. */
int ef_gambler_step(const ef_gambler_params* p, int32_t* state, int32_t* ep_len, float* ep_ret, const int32_t* action,
                    const uint32_t* draw, uint64_t seed, uint64_t t, uint64_t env_offset, int32_t* obs, float* reward,
                    uint8_t* terminated, uint8_t* truncated, int32_t* final_obs, double* stats, uint32_t* err,
                    uint64_t* clock, int64_t n, int32_t autoreset, void* stream);
/* K fused steps (t0 .. t0+K-1), capital in registers, auto-reset always on.  actions [K,n] int32 or NULL (uniform legal
 * bet); rewards [K,n] float / flags [K,n] uint8 (bit0 terminated, bit1 truncated) or NULL. */
int ef_gambler_rollout(const ef_gambler_params* p, int32_t* state, int32_t* ep_len, float* ep_ret, const int32_t* actions,
                       uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K, float* rewards, uint8_t* flags,
                       double* stats, uint32_t* err, uint64_t* clock, int64_t n, void* stream);
int ef_gambler_reset(int32_t start_capital, int32_t* state, int32_t* ep_len, float* ep_ret, int32_t* obs,
                     const uint8_t* mask, int64_t n, void* stream);

typedef struct ef_car_rental_params {
    int32_t max_cars, max_move_cars, rental_credit, move_car_cost;   /* car_rental.py:12-19 */
    int32_t episode_limit;    /* <= 0: none (the reference never ends an episode) */
    int32_t init_equal;       /* reset option: 0 "random" (uniform 0..max_cars per location), 1 "equal" (:134-142) */
    double lam[4];            /* request_lambda[0], request_lambda[1], return_lambda[0], return_lambda[1] */
    double exp_neg_lam[4];    /* exp(-lam[j]) as the host's libm computes it (numpy's random_poisson_mult does the same) */
    /* rates >= 10 draw by numpy's random_poisson_ptrs; its per-rate constants, from the host's libm like numpy's own:
     *   slam = sqrt(lam); b = 0.931 + 2.53*slam; a = -0.059 + 0.02483*b; invalpha = 1.1239 + 1.1328/(b - 3.4);
     *   vr = 0.9277 - 3.6224/(b - 2).   Ignored (may be 0) for rates < 10. */
    double log_lam[4], ptrs_a[4], ptrs_b[4], ptrs_log_invalpha[4], ptrs_vr[4];
} ef_car_rental_params;

/* JacksCarRental.step (:106-121) for n envs: move cars (:54-74), then four legacy-numpy Poisson draws (scipy.stats.
 * poisson.rvs, :89-92) in FP64 over the Philox stream STREAM_ACML of (env, t): the multiplication method for rates < 10,
 * PTRS (transformed rejection with numpy's random_loggam) for rates >= 10, exactly where numpy switches.
 *   state  [n,2] int32 (cars at A, cars at B), 8-byte aligned.   action [n] int32 raw action (0 .. 2*max_move_cars; the
 *   reference does not validate it and neither does the kernel), or NULL: uniform random.
 *   obs / final_obs [n,2]; reward = rentals * credit - |action - max_move| * cost; terminated == 0 always.
 * EF_ERR_UNSUPPORTED for a lam > 1e6. */
int ef_car_rental_step(const ef_car_rental_params* p, int32_t* state, int32_t* ep_len, float* ep_ret,
                       const int32_t* action, uint64_t seed, uint64_t t, uint64_t env_offset, int32_t* obs, float* reward,
                       uint8_t* terminated, uint8_t* truncated, int32_t* final_obs, double* stats, uint64_t* clock,
                       int64_t n, int32_t autoreset, void* stream);
/* K fused steps; actions [K,n] or NULL (uniform random action); rewards [K,n] / flags [K,n] (bit1 truncated) or NULL. */
int ef_car_rental_rollout(const ef_car_rental_params* p, int32_t* state, int32_t* ep_len, float* ep_ret,
                          const int32_t* actions, uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K, float* rewards,
                          uint8_t* flags, double* stats, uint64_t* clock, int64_t n, void* stream);
int ef_car_rental_reset(const ef_car_rental_params* p, int32_t* state, int32_t* ep_len, float* ep_ret, uint64_t seed,
                        uint64_t epoch, uint64_t env_offset, int32_t* obs, const uint8_t* mask, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Labyrinth   (reference: rl_envs_forge/envs/labyrinth/labyrinth.py; maze generation stays on the host)
 * A maze is its finished grid (uint8 [rows*cols], WALL = 0, anything else walkable; constants.py:12-16) plus a start
 * and a target cell (cell = row * cols + col) -- `Labyrinth.maze.{grid, start_position, target_position}`.  n players
 * walk n_mazes mazes: player i uses maze maze_index[i], or (env_offset + i) % n_mazes when maze_index is NULL.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ef_labyrinth_desc {
    int32_t rows, cols, n_mazes;
    int32_t vision_range;            /* state_vision_range (labyrinth.py:20); <= 0: None = the full state matrix */
    int32_t episode_limit;           /* <= 0: none (the reference never truncates) */
    float neutral_reward;            /* reward_schema (:103-108): -0.01 */
    float wall_collision_reward;     /* -1 */
    float target_reached_reward;     /* 10 */
    const uint8_t* grids;            /* device [n_mazes, rows*cols] */
    const int32_t* maze_index;       /* device [n] or NULL */
    const int32_t* start_cell;       /* device [n_mazes] */
    const int32_t* target_cell;      /* device [n_mazes] */
} ef_labyrinth_desc;

/* Labyrinth.step (:245-288) for n players, with same-step auto-reset to the maze's start cell (the reference's own
 * reset() generates a NEW maze, :342-363 -- host work; same-maze reset is its reset(same_seed=True)).
 *   pos      [n] int32 player cell (in/out).   action [n] int32 in {0 UP, 1 RIGHT, 2 DOWN, 3 LEFT}, NULL: random policy.
 *   obs      uint8 [n, rows*cols] (build_state_matrix :210-215: grid, TARGET = 2 at the target, PLAYER = 4 on top) or
 *            [n, (2v+1)^2] (build_state_around_the_player :217-243, WALL outside the maze); after auto-reset.  Nullable.
 *   final_obs  the same view before auto-reset (nullable).  reward: wall_collision on a blocked move (position kept),
 *            target_reached when the move lands on the target (terminated = 1), neutral otherwise.
 *   An action outside 0..3 is the reference's ValueError in Action(action): counted in `err`, player untouched. */
int ef_labyrinth_step(const ef_labyrinth_desc* maze, int32_t* pos, int32_t* ep_len, float* ep_ret, const int32_t* action,
                      uint64_t seed, uint64_t t, uint64_t env_offset, uint8_t* obs, float* reward, uint8_t* terminated,
                      uint8_t* truncated, uint8_t* final_obs, double* stats, uint32_t* err, uint64_t* clock, int64_t n,
                      int32_t autoreset, void* stream);
/* K fused steps without observation rendering (position in registers; render once afterwards with ef_labyrinth_observe).
 * actions [K,n] or NULL (uniform random direction); rewards [K,n] / flags [K,n] (bit0 terminated, bit1 truncated) or NULL. */
int ef_labyrinth_rollout(const ef_labyrinth_desc* maze, int32_t* pos, int32_t* ep_len, float* ep_ret, const int32_t* actions,
                         uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t K, float* rewards, uint8_t* flags,
                         double* stats, uint32_t* err, uint64_t* clock, int64_t n, void* stream);
/* Labyrinth.state (:195-208) for n players: renders `obs` from the current positions, touches nothing else. */
int ef_labyrinth_observe(const ef_labyrinth_desc* maze, const int32_t* pos, uint64_t env_offset, uint8_t* obs, int64_t n,
                         void* stream);
/* Player back to the start cell of its maze, counters cleared; mask nullable. */
int ef_labyrinth_reset(const ef_labyrinth_desc* maze, int32_t* pos, int32_t* ep_len, float* ep_ret, uint64_t env_offset,
                       const uint8_t* mask, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Persistent step server: K GridWorld.step calls (grid_world.py:495-528, with reset :530-551 as auto-reset) of ONE env
 * set WITHOUT a kernel launch per step, for a policy that lives on the device in a kernel of its own.  The server kernel
 * stays resident for all K steps, keeps the (packed) state of its envs in registers and exchanges one CHUNK of
 * ef_gridworld_serve_chunk_envs(n) consecutive envs per CTA with the action producer through two sequence words per chunk:
 *     producer:  (k > 0: wait obs_seq[c] >= k)   read the outputs of step k-1   write action[chunk c]   act_seq[c] = k + 1
 *     server  :  wait act_seq[c] >= k + 1         read action[chunk c]            step, write the outputs  obs_seq[c] = k + 1
 * with release stores / acquire loads at gpu scope (`st.release.gpu` / `ld.acquire.gpu`, or __threadfence() + volatile
 * accesses) and L2-coherent reads of whatever the other kernel wrote (`ld.global.cg` / __ldcg): both kernels run at the
 * same time, on different streams, and must be resident together -- the server occupies one CTA of 256 threads and at
 * most 96 registers per chunk, at most EF_SERVE_CHUNKS_PER_SM chunks per SM (every chunk is a dependency chain of its
 * own, and the chains of an SM hide each other's latencies).  Results are bit-identical to K ef_gridworld_step calls with the
 * same actions.  Packed state layout only (`pos` holds cell | counter << 16, see ef_gridworld_step), n a multiple of 4,
 * table small enough for shared memory AND for all chunks' CTAs to be resident at once next to the kernel's own 17 KB of
 * shared memory (else EF_ERR_UNSUPPORTED: a chunk waiting for an SM slot would wait for CTAs that wait for their actions).
 * While a CTA waits for a step's actions it computes the step's Philox draws.  act_seq / obs_seq: [ceil(n / chunk_envs)] uint32, zeroed by the caller before
 * both kernels are launched.  The producer kernel must already be LOADED when the server starts (launched once before, or
 * cudaFuncGetAttributes on it): CUDA loads kernels lazily and a first launch may synchronise the context, i.e. wait for
 * the server that is waiting for it.  A wait that is not satisfied within ~5 s gives up, sets status[0] = 1 (nullable) and ends
 * the kernel after writing the state back; `clock` / `t0` advance by the steps actually taken.
 * ef_policy_demo_serve is a stand-in producer (tests, bench): action = (row * 3 + col + k) & 3 of the observation the env
 * returned for step k - 1 (the content of `obs` at launch for k = 0).
 * ------------------------------------------------------------------------------------------------------------ */
#ifndef EF_SERVE_CHUNKS_PER_SM
#define EF_SERVE_CHUNKS_PER_SM 2
#endif
int ef_gridworld_serve_chunk_envs(int64_t n);   /* envs per chunk for a batch of n (<= 0: the batch does not fit one CTA per SM) */
int ef_gridworld_serve(const ef_grid_desc* grid, int32_t* pos, float* ep_ret, const int32_t* action, uint64_t seed,
                       uint64_t t0, uint64_t env_offset, int32_t K, int32_t* obs, float* reward, uint8_t* terminated,
                       uint8_t* truncated, double* stats, uint32_t* err, uint64_t* clock, uint32_t* act_seq,
                       uint32_t* obs_seq, uint32_t* status, int64_t n, void* stream);
int ef_policy_demo_serve(const int32_t* obs, int32_t* action, const uint32_t* obs_seq, uint32_t* act_seq, int32_t K,
                         uint32_t* status, int64_t n, void* stream);
/* The same stand-in policy as an ordinary kernel, one launch per step (the baseline the server is measured against). */
int ef_policy_demo_step(const int32_t* obs, int32_t* action, int32_t k, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Episode statistics: reduce per-env running (ep_len, ep_ret) of the envs still in flight into `out[0..3]`
 * (count, sum len, sum ret, sum ret^2) -- complements `stats`, which only sees finished episodes.
 * ------------------------------------------------------------------------------------------------------------ */
int ef_episode_stat_reduce(const int32_t* ep_len, const float* ep_ret, double* out, int64_t n, void* stream);

/* Fold the replicated `stats` accumulators of up to EF_STATS_MAX_SETS env sets into ONE vector:
 *     out[f] = sum over sets s, replicas r of raw_sets[s][r * EF_STATS_STRIDE + f],   f < EF_STATS_LEN,
 * in a fixed order (bit-reproducible).  `raw_sets` is a HOST array of n_sets DEVICE pointers (copied into the kernel
 * parameters); `out` is EF_STATS_LEN device doubles.  One tiny CTA, stream-ordered, capturable in a CUDA graph -- this is
 * what feeds the cross-GPU sum of the episode statistics (the engine's only collective). */
#define EF_STATS_MAX_SETS 32
int ef_stats_fold(const double* const* raw_sets, int32_t n_sets, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Cross-GPU sum of the statistics vector over NVLink peer memory, fused with the fold above (one kernel: fold the
 * local replicas, store the 64-byte result into every peer's mailbox, wait for the peers' vectors, add them in rank
 * order).  One process per GPU; every rank creates a comm, the 64-byte IPC handles are exchanged by the caller
 * (torch.distributed all_gather in the shipped host code), then every rank connects.  The mailbox (world * 2 slots of
 * 128 bytes) is the ONE piece of device memory this library allocates itself: it must be a cudaMalloc allocation of
 * its own to be exportable with cudaIpcGetMemHandle.
 *   ef_stats_fold_allreduce  out_local (nullable) receives this rank's folded vector, out_global the sum over ranks --
 *                            bit-identical on every rank (same order of additions).  n_sets == 0 (raw_sets ignored): the
 *                            vector was folded earlier by ef_stats_fold and out_local is the INPUT; this is how the
 *                            exchange runs on a side stream while the stepping stream goes on.  Collective: every rank must launch it
 *                            the same number of times.  A peer that does not show up within ~10 s makes the kernel give up
 *                            and set status[0] (device uint32, nullable) to 1 instead of hanging the GPU.
 * ------------------------------------------------------------------------------------------------------------ */
#define EF_PEER_HANDLE_BYTES 64
typedef struct ef_peer_comm ef_peer_comm;
int ef_peer_comm_create(int32_t rank, int32_t world, ef_peer_comm** comm, uint8_t* handle_out);
int ef_peer_comm_connect(ef_peer_comm* comm, const uint8_t* all_handles);   /* world * EF_PEER_HANDLE_BYTES, rank-major */
int ef_peer_comm_destroy(ef_peer_comm* comm);
int ef_stats_fold_allreduce(ef_peer_comm* comm, const double* const* raw_sets, int32_t n_sets, double* out_local,
                            double* out_global, uint32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host-buffer step: the call a host-side actor makes once per step (the reference's `env.step(action)` with the
 * action and the results in HOST memory: grid_world.py:495, cart_pole.py:107, pendulum_disk.py:108,
 * k_armed_bandit.py:175).  Same kernels and same results as the device-buffer entry points above; what is added is
 * the PCIe plumbing.  The batch is cut into `chunks` contiguous index ranges (starts are multiples of 1024 envs, global
 * env index = env_offset + start, so chunking never changes a result) and pipelined:
 *     actions of chunk c+1 cross PCIe host->device  ||  kernel of chunk c+1  ||  outputs of chunk c cross device->host
 *   h_*   PINNED (page-locked, UVA-addressable) host buffers of the whole batch.  h_action is read, the others written.
 *   d_*   device buffers of the whole batch the kernels write before the copy engine drains them (caller-owned
 *         scratch).  An output is skipped when either its d_ or h_ pointer is NULL.
 *   d_action  device staging buffer [n] or NULL.  NULL: the step kernel reads the actions directly from h_action
 *         (zero-copy over PCIe); non-NULL: each chunk's actions are first copied by the copy engine.
 *   The call is asynchronous; when `stream` has been synchronised the h_ outputs are valid.  There is no `clock`
 *   argument: a host-buffer step is not graph-replayable (its `t` is a plain argument).
 * ef_host_pipe holds two copy streams and the events that order them; it owns no device memory.  Create it once per
 * env set on the device it is used on; not thread-safe (one in-flight step per pipe).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct ef_host_pipe ef_host_pipe;
int ef_host_pipe_create(ef_host_pipe** pipe);
int ef_host_pipe_destroy(ef_host_pipe* pipe);

int ef_gridworld_step_host(ef_host_pipe* pipe, const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                           const int32_t* h_action, int32_t* d_action, uint64_t seed, uint64_t t, uint64_t env_offset,
                           int32_t* d_obs, float* d_reward, uint8_t* d_terminated, uint8_t* d_truncated,
                           int32_t* d_final_obs, int32_t* h_obs, float* h_reward, uint8_t* h_terminated,
                           uint8_t* h_truncated, int32_t* h_final_obs, double* stats, uint32_t* err, int64_t n,
                           int32_t autoreset, int32_t chunks, void* stream);
/* The same pipeline over the narrow wire format (see ef_gridworld_step_u8): uint8 actions, uint8 [n,2] observations. */
int ef_gridworld_step_u8_host(ef_host_pipe* pipe, const ef_grid_desc* grid, int32_t* pos, int32_t* ep_len, float* ep_ret,
                              const uint8_t* h_action, uint8_t* d_action, uint64_t seed, uint64_t t, uint64_t env_offset,
                              uint8_t* d_obs, float* d_reward, uint8_t* d_terminated, uint8_t* d_truncated,
                              uint8_t* d_final_obs, uint8_t* h_obs, float* h_reward, uint8_t* h_terminated,
                              uint8_t* h_truncated, uint8_t* h_final_obs, double* stats, uint32_t* err, int64_t n,
                              int32_t autoreset, int32_t chunks, void* stream);
int ef_cartpole_step_host(ef_host_pipe* pipe, const ef_pendulum_params* p, float* state, int32_t* ep_len, float* ep_ret,
                          const float* h_action, float* d_action, uint64_t seed, uint64_t t, uint64_t env_offset,
                          float* d_obs, float* d_reward, uint8_t* d_terminated, uint8_t* d_truncated, float* d_final_obs,
                          float* d_acc, float* h_obs, float* h_reward, uint8_t* h_terminated, uint8_t* h_truncated,
                          float* h_final_obs, float* h_acc, double* stats, uint32_t* err, int64_t n, int32_t autoreset, int32_t chunks,
                          void* stream);
int ef_pendulum_disk_step_host(ef_host_pipe* pipe, const ef_pendulum_params* p, float* state, int32_t* ep_len,
                               float* ep_ret, const float* h_action, float* d_action, uint64_t seed, uint64_t t,
                               uint64_t env_offset, float* d_obs, float* d_reward, uint8_t* d_terminated,
                               uint8_t* d_truncated, float* d_final_obs, float* d_acc, float* h_obs, float* h_reward,
                               uint8_t* h_terminated, uint8_t* h_truncated, float* h_final_obs, float* h_acc,
                               double* stats, uint32_t* err, int64_t n, int32_t autoreset, int32_t chunks, void* stream);
int ef_bandit_step_host(ef_host_pipe* pipe, const ef_arm* arms, int32_t k, const int32_t* h_action, int32_t* d_action,
                        uint64_t seed, uint64_t t0, uint64_t env_offset, int32_t* d_obs, float* d_reward, int32_t* h_obs,
                        float* h_reward, double* stats, uint32_t* err, int64_t n, int32_t chunks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ENVFORGE_B200_H */
