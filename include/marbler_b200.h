/* marbler_b200 -- C ABI of the batched MARBLER environment step on B200 (sm_100a).
 *
 * The reference (GT-STAR-Lab/MARBLER) is pure Python and has NO FFI of its own for this path; the
 * only native boundary it crosses is cvxopt's CPython extension.  The entry points
 * below are therefore the binding surface a maintainer would add (INTEGRATION.md shows the ctypes
 * stub); each cites the reference interface it replaces.  Conventions: extern "C", plain pointers
 * and sizes, no C++/torch types, no exceptions; 0 = ok, <0 = error (text via mrb_last_error).
 * Stream-ordered: every call enqueues on the caller's CUDA stream and (except *_host) never
 * synchronises.  The library never allocates caller-visible buffers: state / output buffers are
 * device memory owned by the caller (PyTorch tensors) and handed over with mrb_bind.
 * One handle per (process, GPU); a handle is not thread-safe (the reference is single-threaded).
 */
#ifndef MARBLER_B200_H
#define MARBLER_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MRB_ABI_VERSION 2
#define MRB_MAX_ROBOTS 32
#define MRB_MAX_PREY 32

/* scenario ids: robotarium_gym/wrapper.py:12-16 env_dict, robotarium_gym/__init__.py:4-10 */
enum { MRB_PCP = 0, MRB_WAREHOUSE = 1, MRB_MATERIAL = 2, MRB_ARCTIC = 3, MRB_SIMPLE = 4 };
/* message codes: roboEnv.py:85-90 '' / 'collision' / 'boundary' / 'collision_boundary' */
enum { MRB_MSG_NONE = 0, MRB_MSG_COLLISION = 1, MRB_MSG_BOUNDARY = 2, MRB_MSG_COLLISION_BOUNDARY = 3 };
enum { MRB_OK = 0, MRB_E_ARG = -1, MRB_E_CUDA = -2, MRB_E_UNSUPPORTED = -3, MRB_E_STATE = -4 };

/* grid spawn sampler = rps generate_initial_conditions + utilities/misc.py:49-63
 * generate_initial_locations:  x = ((ix*spacing - w2) + sx1) + sx2,  y = ((iy*spacing - h2) + sy1) + sy2,
 * (ix, iy) = divmod(cell, yr), `count` distinct cells of the xr*yr grid. */
typedef struct mrb_spawn {
    int32_t count, xr, yr, random_theta;
    double spacing, w2, h2, sx1, sx2, sy1, sy2;
} mrb_spawn;

/* The scenario config.yaml keys the step path reads, already resolved. */
typedef struct mrb_config {
    int32_t struct_size;            /* sizeof(mrb_config), checked by mrb_create */
    int32_t scenario;               /* MRB_PCP ... */
    int32_t num_robots;             /* PCP: predator+capture (PredatorCapturePrey.py:19); else n_agents; 2..32 */
    int32_t update_frequency;       /* roboEnv.py:52 */
    int32_t ctrl_period;            /* 15, roboEnv.py:63 */
    int32_t robotarium;             /* roboEnv.py:63: controller every sub-step */
    int32_t penalize_violations;    /* roboEnv.py:82 */
    int32_t barrier_default;        /* 0 'safe' (certificate2, r=.2), 1 'default' (r=.17): controller.py:13-16 */
    int32_t max_episode_steps;
    int32_t num_neighbors;          /* PCP / Warehouse */
    int32_t capability_aware;       /* PCP / MaterialTransport */
    int32_t num_prey, num_predators;                 /* PCP */
    int32_t n_fast, small_torque, large_torque;      /* MaterialTransport */
    int32_t auto_reset;             /* re-sample finished envs at the end of mrb_step */
    int32_t track_dist;             /* info['dist_travelled'] (roboEnv.py:55-56,93) */
    int32_t collect_stats;          /* episode statistics vector, see mrb_buffers.stats */
    int32_t reserved0;
    double left, right, up, down;   /* LEFT/RIGHT/UP/DOWN */
    double step_dist;               /* step_dist; ArcticTransport normal_step */
    double fast_step, slow_step;    /* MaterialTransport / ArcticTransport */
    double predator_radius, capture_radius;
    double time_penalty, sense_reward, capture_reward;
    double load_reward, unload_reward, goal_width;   /* Warehouse; MT: load/unload multiplier, end_goal_width */
    double zone1_radius;
    double not_reached_penalty, dist_multiplier;     /* ArcticTransport */
    double reward_scaler;                            /* Simple */
    double violation_reward;        /* -5 / -5 / -6 / -30 / -5 (literals in each scenario's step()) */
    double zone_mu[2], zone_sigma[2];                /* MaterialTransport zone1/zone2 normal(loc, scale) */
    mrb_spawn spawn_robots, spawn_other;             /* robots; PCP prey / Simple goal */
    /* rps RobotariumABC._validate collision test (roboEnv.py:80-94 reads its counters): a pair collides when
     * |(p_j + collision_offset (cos th_j, sin th_j)) - (p_k + collision_offset (cos th_k, sin th_k))| <= collision_diameter.
     * rps is not vendored in the reference and its pinned commit (6bb184e) cannot be read in the build container,
     * so both published forms are supported: centre to centre (collision_offset = 0; the SURVEY.md App. A.4
     * restatement, the default of the Python layer) and heading-projected points (collision_offset = 0.025). */
    double collision_diameter, collision_offset;
} mrb_config;

/* Caller-owned device buffers.  State is structure-of-arrays with the env index fastest:
 *   state_f64[row][env], state_i32[row][env]   (row counts from mrb_state_rows; rows named in
 *   marbler_b200/layout.py and DESIGN.md).  Outputs:
 *   obs f32 [B][N][D] | reward f32 [B][N] | done u8 [B] | message u8 [B] | remaining i32 [B] |
 *   dist f32 [B][N] (may be NULL when track_dist == 0) | stats f64 [MRB_NUM_STATS] (may be NULL). */
#define MRB_NUM_STATS 16
enum { MRB_STAT_EPISODES = 0, MRB_STAT_RETURN = 1, MRB_STAT_LENGTH = 2, MRB_STAT_COLLISION = 3,
       MRB_STAT_BOUNDARY = 4, MRB_STAT_SCENARIO = 5, MRB_STAT_ENV_STEPS = 6, MRB_STAT_QP_SOLVES = 7,
       MRB_STAT_QP_ITERS = 8, MRB_STAT_TIMEOUTS = 9,
       MRB_STAT_QP_STALLS = 10, /* solves that ran >= 25 iterations: cvxopt's iteration caught in a limit cycle */
       MRB_STAT_SUBSTEPS = 11,  /* simulator sub-steps executed (update_frequency per env step unless a violation ends it early) */
       /* iterations the hardware ran for the solves: with one env per thread a warp iterates until its slowest
        * env has converged, so this is the sum over solves of (warp maximum x lanes of the warp); divided by
        * MRB_STAT_QP_ITERS it is the divergence overhead of the solver loop (1.0 for the one-env-per-warp kernels) */
       MRB_STAT_QP_ITERS_WARP = 12 };
typedef struct mrb_buffers {
    double *state_f64;
    int32_t *state_i32;
    float *obs;
    float *reward;
    uint8_t *done;
    uint8_t *message;
    int32_t *remaining;
    float *dist;
    double *stats;
    /* optional float64 copies of obs / reward (NULL = off).  The reference returns float64 numpy observations and
     * Python-float rewards (e.g. PredatorCapturePrey.py:176); the single-env drop-in path binds these so that its
     * return values are not rounded through float32. */
    double *obs_f64;    /* [B][N][D] */
    double *reward_f64; /* [B][N] */
} mrb_buffers;

typedef struct mrb_env mrb_env;

int mrb_version(void);
/* replaces Wrapper.__init__ -> env_dict[env_name](args) (wrapper.py:20-34): builds the handle for
 * num_envs independent envs whose global ids are env_id0 .. env_id0+num_envs-1 (RNG stream offset
 * of this rank when envs are sharded over GPUs). */
int mrb_create(const mrb_config *cfg, int device, int64_t num_envs, int64_t env_id0, mrb_env **out);
int mrb_destroy(mrb_env *env);
const char *mrb_last_error(const mrb_env *env);          /* env may be NULL: last create error */
/* layout queries (spaces: <Scenario>.get_observation_space / get_action_space) */
int mrb_state_rows(const mrb_env *env, int32_t *rows_f64, int32_t *rows_i32);
int mrb_obs_dim(const mrb_env *env);
int mrb_num_actions(const mrb_env *env);
int mrb_bind(mrb_env *env, const mrb_buffers *buffers);
/* replaces <Scenario>.reset() + roboEnv.reset() (e.g. PredatorCapturePrey.py:114-136,
 * roboEnv.py:27-36,98-118).  mask: device u8[B] (NULL = all envs).  Zeroes obs of the reset envs
 * (the reference returns an all-zero observation from reset). */
int mrb_reset(mrb_env *env, const uint8_t *mask, uint64_t seed, void *cuda_stream);
/* replaces Wrapper.step(action_n) (wrapper.py:41-44) -> <Scenario>.step -> roboEnv.step
 * (roboEnv.py:38-96) -> Controller.set_velocities (controller.py:20-25) -> rps / cvxopt.
 * actions: device i32 [B][N]. */
int mrb_step(mrb_env *env, const int32_t *actions, void *cuda_stream);
/* same step with HOST buffers (pinned for full speed; pageable works): H2D actions, step, D2H
 * obs/reward/done/message, then synchronises.  Batches of >= 16,384 envs are cut into up to 8 chunks,
 * each on its own library-internal stream (ordered after the caller's stream), so that the PCIe copies
 * of one chunk overlap the kernels of the others.  NULL host outputs are skipped.  When every buffer
 * is pinned (cudaHostAlloc / cudaHostRegister) the kernel stores reward / done / message directly into
 * the host buffers through their device aliases and only obs is downloaded by the copy engine;
 * environment MRB_HOST_DIRECT=1 sends obs the same way, =0 downloads everything.  The device-side
 * buffers bound with mrb_bind are written in every mode. */
int mrb_step_host(mrb_env *env, const int32_t *actions_host, float *obs_host, float *reward_host,
                  uint8_t *done_host, uint8_t *message_host, void *cuda_stream);
/* ---- state access by field (checkpoint / restore, parity injection).  The reference keeps this state in Python
 * attributes of the scenario, its agents and roboEnv (cited per field); here it is packed into the SoA rows of
 * state_f64 / state_i32, and these two calls convert between the rows and plain per-field HOST arrays for the envs
 * [env_lo, env_lo + count).  NULL fields are skipped (get) / left unchanged (set); fields of other scenarios are
 * ignored.  Both calls are ordered after the work already enqueued on cuda_stream and synchronise before returning. */
typedef struct mrb_state_fields {
    int32_t struct_size;         /* sizeof(mrb_state_fields) */
    int32_t reserved0;
    double *poses;               /* [count][3][N]  scenario.agent_poses = robotarium.poses (x row, y row, theta row) */
    double *prev_pose;           /* [count][3][N]  roboEnv.previous_pose (roboEnv.py:55-59); only x, y are state, theta reads 0 */
    double *episode_return;      /* [count]        sum of team rewards of the running episode (statistics only) */
    int32_t *episode_steps;      /* [count]        scenario.episode_steps */
    int32_t *prev_valid;         /* [count]        roboEnv.previous_pose is not None (roboEnv.py:117) */
    int32_t *episode_count;      /* [count]        episodes started so far: Philox counter of the next auto-reset */
    double *prey_loc;            /* [count][P][2]  PCP prey_loc (PredatorCapturePrey.py:128-132) */
    uint8_t *prey_sensed;        /* [count][P]     PCP prey_sensed */
    uint8_t *prey_captured;      /* [count][P]     PCP prey_captured */
    uint8_t *loaded;             /* [count][N]     Warehouse agent.loaded (warehouse.py:145-178) */
    int32_t *load;               /* [count][N]     MaterialTransport agent.load */
    int32_t *zone_load;          /* [count][2]     MaterialTransport zone1_load, zone2_load (MaterialTransport.py:99-100) */
    int32_t *messages;           /* [count][4]     MaterialTransport messages (MaterialTransport.py:119-120) */
    uint8_t *grid;               /* [count][8][12] ArcticTransport grid (ArcticTransport.py:56-82) */
    int32_t *goal_col;           /* [count]        ArcticTransport goal_loc[1] */
    int32_t *pixel_type;         /* [count][N]     ArcticTransport agent.pixel_type (agent.py:37-39) */
    uint8_t *reached_goal;       /* [count][N]     ArcticTransport agent.reached_goal */
    double *goal;                /* [count][2]     Simple goal_loc (simple.py:141-144) */
} mrb_state_fields;
int mrb_get_state(mrb_env *env, int64_t env_lo, int64_t count, const mrb_state_fields *out, void *cuda_stream);
int mrb_set_state(mrb_env *env, int64_t env_lo, int64_t count, const mrb_state_fields *in, void *cuda_stream);

/* unit entry for the barrier-certificate QP alone (rps create_single_integrator_barrier_certificate{,2}
 * -> cvxopt.solvers.qp; called at controller.py:23).  Device SoA: dxi, xi, u are f64 [2][N][B];
 * iters i32 [B] (may be NULL). */
int mrb_barrier_qp(int device, int32_t num_robots, int32_t barrier_default, int64_t num_problems,
                   const double *dxi, const double *xi, double *u, int32_t *iters, void *cuda_stream);

/* ---- on-device policy: the reference's evaluation loop utilities/misc.py:134-221
 * (run_env) calls utilities/rnn_agent.py:5-29 RNNAgent / utilities/rnn_ns_agent.py:5-36 RNNNSAgent on the
 * host once per env step: q, h = model(obs, h); actions = argmax(q).  mrb_policy_act does that for every
 * agent of every env in one kernel: fc1 -> ReLU -> GRUCell (or Linear+ReLU when use_rnn == 0) -> fc2 ->
 * greedy argmax, tensor-core MMAs on FP16 operands with FP32 accumulation and FP32 gates (tcgen05 / TMEM for
 * hidden 128 + GRUCell with <= 8 actions, mma.sync otherwise) - or, with desc.accurate, in float32 throughout -,
 * reading the env's obs buffer and writing the actions buffer mrb_step consumes (no host round trip). */
typedef struct mrb_policy_desc {
    int32_t struct_size;        /* sizeof(mrb_policy_desc): ABI check */
    int32_t obs_dim;            /* D: width of one agent's row in the obs buffer */
    int32_t input_dim;          /* fc1.in_features = D (+ n_agents when obs_agent_id, misc.py:161-162) */
    int32_t hidden_dim;         /* 64 or 128 (model json "hidden_dim") */
    int32_t n_actions;          /* <= 24 */
    int32_t n_agents;
    int32_t obs_agent_id;       /* append the one-hot agent id to the observation */
    int32_t use_rnn;            /* 1: nn.GRUCell, 0: nn.Linear + ReLU (rnn_agent.py:11-14) */
    int32_t non_shared;         /* 1: one weight set per agent (RNNNSAgent), 0: one shared set */
    int32_t accurate;           /* 1: float32 operands and accumulation like the reference (plain FMA kernel, ~10x the
                                 * time): the greedy actions of a checkpoint are the reference's; 0: FP16 tensor-core
                                 * operands (q within ~1e-3 |q|, about one near-tie decision in a thousand differs) */
} mrb_policy_desc;
typedef struct mrb_policy mrb_policy;
/* weights: HOST float32, (non_shared ? n_agents : 1) sets back to back, each set in state_dict order and
 * torch layout ([out][in] row-major): fc1.weight, fc1.bias, then rnn.weight_ih [3H][H], rnn.weight_hh [3H][H],
 * rnn.bias_ih, rnn.bias_hh (use_rnn) or rnn.weight [H][H], rnn.bias (otherwise), then fc2.weight, fc2.bias. */
int mrb_policy_create(const mrb_policy_desc *desc, int device, const float *weights, int64_t num_weights,
                      mrb_policy **out);
int mrb_policy_destroy(mrb_policy *policy);
const char *mrb_policy_last_error(const mrb_policy *policy);
/* obs f32 [B][N][D], hidden f32 [B][N][H] (in/out), actions i32 [B][N] (out), q f32 [B][N][n_actions] (out,
 * may be NULL), fresh u8 [B] (may be NULL): envs flagged != 0 start a new episode - their hidden state and
 * observation are taken as zero (run_env re-zeroes hs and feeds reset()'s all-zero obs, misc.py:156,219).
 * All device pointers; stream-ordered, no synchronisation. */
int mrb_policy_act(mrb_policy *policy, int64_t num_envs, const float *obs, float *hidden, int32_t *actions,
                   float *q, const uint8_t *fresh, void *cuda_stream);

/* FP64 roofline denominator, measured: runs a kernel of independent DFMA chains (8 per thread, one 1,024-thread
 * CTA per SM) for about `milliseconds` on `device`, timed with CUDA events, and returns the rate of the best launch
 * in TFLOP/s (2 flops per DFMA).  The env step is bound by FP64 issue, not by HBM, so bench.py
 * reports the step kernels' FP64 flop rate against this number. */
int mrb_fp64_peak(int device, double milliseconds, double *tflops);

/* number of kernels this library has launched in the process so far (bench.py's gpu_launches) */
int64_t mrb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
