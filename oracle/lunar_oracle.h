/* lunar_oracle.h -- CPU ORACLE for LunarLander-v2 (test infrastructure, NOT product code).
 * See lunar_oracle.c.  PARITY UNPINNED: Box2D (box2d-py 2.3.5) is not available here. */
#ifndef LUNAR_ORACLE_H
#define LUNAR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct orc_lunar orc_lunar;
orc_lunar *orc_lunar_create(int64_t n, int max_episode_steps);
orc_lunar *orc_lunar_create_ex(int64_t n, int max_episode_steps, int continuous, int enable_wind, double gravity,
                               double wind_power, double turbulence_power);
void orc_lunar_set_wind_idx(orc_lunar *v, const int32_t *wind_idx, const int32_t *torque_idx);
void orc_lunar_get_wind_idx(const orc_lunar *v, int32_t *wind_idx, int32_t *torque_idx);
void orc_lunar_step_cont(orc_lunar *v, const float *actions, float *obs, double *reward, uint8_t *terminated,
                         uint8_t *truncated, float *final_obs);
void orc_lunar_engines(int continuous, int action, const float *caction, double ang, double posx, double posy,
                       double disp0, double disp1, float out[8], double cost[2], int32_t on[2]);
void orc_lunar_destroy(orc_lunar *v);
void orc_lunar_seed_range(orc_lunar *v, const uint32_t base[4], int64_t first);
void orc_lunar_reset(orc_lunar *v, float *obs);
int64_t orc_lunar_step(orc_lunar *v, const int64_t *actions, float *obs, double *reward, uint8_t *terminated,
                       uint8_t *truncated, float *final_obs);
void orc_lunar_get_bodies(const orc_lunar *v, int64_t i, float out[18], int32_t flags[6]);
int64_t orc_lunar_step_mt(orc_lunar *v, const int64_t *actions, float *obs, double *reward, uint8_t *terminated,
                          uint8_t *truncated, float *final_obs, int nthreads);
void orc_lunar_step_cont_mt(orc_lunar *v, const float *actions, float *obs, double *reward, uint8_t *terminated,
                            uint8_t *truncated, float *final_obs, int nthreads);
void orc_lunar_get_terrain(const orc_lunar *v, int64_t i, float *y11);
void orc_lunar_get_stats(const orc_lunar *v, int32_t *out);
void orc_lunar_toi_stats(const orc_lunar *v, int64_t out[3]);
void orc_lunar_set_toi(int on);
int orc_b2l_toi_probe(const float *poly_xy, int n, const float c0[2], float a0, const float c1[2], float a1,
                      const float v1[2], const float v2_[2], float *t_out);
float orc_b2l_distance_probe(const float *poly_xy, int n, const float c[2], float a, const float v1[2], const float v2_[2],
                             int *cache_count);
void orc_lunar_set_body_velocity(orc_lunar *v, int64_t i, int body, float vx, float vy, float w);
#ifdef __cplusplus
}
#endif
#endif
