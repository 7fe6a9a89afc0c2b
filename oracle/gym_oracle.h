/*
 * gym_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the step()/reset() hot path of openai/gym 0.26.2
 * (classic control + TimeLimit + SyncVectorEnv autoreset + numpy PCG64 /
 * SeedSequence).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; nothing under
 * gym_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit against the
 * reference itself (imported from /root/reference in the build container by
 * oracle/gen_golden.py) and against the committed fixtures in tests/golden/.
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef GYM_ORACLE_H
#define GYM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* env kinds (same numbering as include/b200gym.h, restated independently) */
enum {
    ORC_CARTPOLE = 0,            /* gym/envs/classic_control/cartpole.py            */
    ORC_MOUNTAINCAR = 1,         /* gym/envs/classic_control/mountain_car.py        */
    ORC_MOUNTAINCAR_CONT = 2,    /* gym/envs/classic_control/continuous_mountain_car.py */
    ORC_PENDULUM = 3,            /* gym/envs/classic_control/pendulum.py            */
    ORC_ACROBOT = 4,             /* gym/envs/classic_control/acrobot.py             */
    ORC_NUM_KINDS = 5
};

typedef struct orc_vec orc_vec;

/* dims of the per-env observation / action / state vectors for a kind */
int orc_obs_dim(int kind);
int orc_act_dim(int kind);      /* 0 => Discrete (int64 actions), k>0 => Box(k) float32 */
int orc_state_dim(int kind);
int orc_num_actions(int kind);  /* Discrete n, 0 for Box */

/*
 * Vector of n independent envs of one kind, with the wrapper stack that
 * gym.make applies fused in (TimeLimit: gym/wrappers/time_limit.py:39-68) and
 * SyncVectorEnv's same-step autoreset (gym/vector/sync_vector_env.py:135-169).
 * max_episode_steps <= 0 means "no TimeLimit".
 * param0: Pendulum g (default 10.0), MountainCar/MountainCarContinuous
 *         goal_velocity (default 0.0); ignored for the others.
 */
orc_vec *orc_vec_create(int kind, int64_t n, int max_episode_steps, double param0);
void orc_vec_destroy(orc_vec *v);

/* numpy SeedSequence(seed).generate_state(4, uint64) for seed < 2^128 given as
 * four little-endian uint32 words (gym/utils/seeding.py:24-26). */
void orc_seed_sequence(const uint32_t ent[4], uint64_t out[4]);

/* Seed env i with Generator(PCG64(SeedSequence(seed))) where seed is the
 * 128-bit integer ent[0..3] (little-endian uint32 words). */
void orc_vec_seed_env(orc_vec *v, int64_t i, const uint32_t ent[4]);
/* SyncVectorEnv.reset_wait int-seed fan-out: env i <- base + first + i
 * (gym/vector/sync_vector_env.py:106-107). */
void orc_vec_seed_range(orc_vec *v, const uint32_t base[4], int64_t first);

/* raw PCG64 state access: 4 uint64 per env {state_hi, state_lo, inc_hi, inc_lo} */
void orc_vec_get_rng(const orc_vec *v, uint64_t *out);
void orc_vec_set_rng(orc_vec *v, const uint64_t *in);
/* next double of env i's stream (for KATs) */
double orc_vec_next_double(orc_vec *v, int64_t i);

/*
 * reset(): draws a fresh initial state for every env whose mask byte is
 * non-zero (all envs when mask == NULL) from that env's PCG64 stream, zeroes
 * its TimeLimit counter and writes its float32 observation row.
 * bounds: NULL for the defaults; otherwise {low, high} (CartPole, Acrobot,
 * MountainCar*: options["low"/"high"], classic_control/utils.py:17-46) or
 * {x_init, y_init} (Pendulum: pendulum.py:141-159).
 */
void orc_vec_reset(orc_vec *v, const uint8_t *mask, const double *bounds, float *obs);

/*
 * step(): one SyncVectorEnv.step_wait.  actions: int64[n] for Discrete kinds,
 * float32[n*act_dim] for Box kinds.  Outputs: obs float32[n*obs_dim] (the
 * post-autoreset observation), reward float64[n], terminated/truncated
 * uint8[n], final_obs float32[n*obs_dim] (rows valid where
 * terminated|truncated).  Returns the number of invalid Discrete actions seen
 * (the reference raises AssertionError for those; such envs are left
 * untouched).  nthreads > 1 splits the env range over that many pthreads.
 */
int64_t orc_vec_step(orc_vec *v, const void *actions, float *obs, double *reward,
                     uint8_t *terminated, uint8_t *truncated, float *final_obs,
                     int nthreads);

/* float64 state access [n][state_dim] (+ the elapsed-step counters) */
void orc_vec_get_state(const orc_vec *v, double *state, int32_t *elapsed);
void orc_vec_set_state(orc_vec *v, const double *state, const int32_t *elapsed);

#ifdef __cplusplus
}
#endif
#endif
