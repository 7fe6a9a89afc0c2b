/* walker_oracle.h -- CPU ORACLE for BipedalWalker-v3 / BipedalWalkerHardcore-v3 (test infrastructure, NOT product code).
 * See walker_oracle.c.  PARITY UNPINNED for the Box2D arithmetic; the numpy RNG side is pinned. */
#ifndef WALKER_ORACLE_H
#define WALKER_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_walker orc_walker;
orc_walker *orc_walker_create(int64_t n, int max_episode_steps);
orc_walker *orc_walker_create_ex(int64_t n, int max_episode_steps, int hardcore);
int orc_walker_get_polys(const orc_walker *v, int64_t i, float *out);
void orc_walker_destroy(orc_walker *v);
void orc_walker_seed_range(orc_walker *v, const uint32_t base[4], int64_t first);
void orc_walker_reset(orc_walker *v, float *obs);
void orc_walker_step(orc_walker *v, const float *actions, float *obs, double *reward, uint8_t *terminated,
                     uint8_t *truncated, float *final_obs);
void orc_walker_get_terrain(const orc_walker *v, int64_t i, float *y200);
void orc_walker_get_bodies(const orc_walker *v, int64_t i, float out[30], int32_t flags[4]);
void orc_rng_sequence(const uint32_t ent[4], const int32_t *ops, int64_t n, double *out);
void orc_walker_step_mt(orc_walker *v, const float *actions, float *obs, double *reward, uint8_t *terminated,
                        uint8_t *truncated, float *final_obs, int nthreads);
void orc_walker_action_flow(double shaping_delta, const float action[4], float motor_speed[4], float max_torque[4],
                            double *reward);
void orc_walker_get_stats(const orc_walker *v, int32_t *out);
void orc_walker_toi_stats(const orc_walker *v, int64_t out[3]);
void orc_walker_set_toi(int on);
#ifdef __cplusplus
}
#endif
#endif
