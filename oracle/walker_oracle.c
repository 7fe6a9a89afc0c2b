/*
 * walker_oracle.c -- CPU ORACLE for BipedalWalker-v3 (test infrastructure, NOT product code).
 *
 * Restates gym/envs/box2d/bipedal_walker.py (reset :425-515, _generate_terrain :277-402,
 * _generate_clouds :404-423, step :517-606, ContactDetector :80-98, LidarCallback :504-510) on top
 * of oracle/b2lite.h, the from-scratch restatement of the Box2D 2.3 subset the env exercises.
 * Both terrains: BipedalWalker-v3 and BipedalWalkerHardcore-v3 (hardcore=True: stumps, stairs, pits as static
 * box polygons on top of the edge chain, :300-373).
 *
 * PARITY UNPINNED for the rigid-body arithmetic (no Box2D here, see b2lite.h).  The numpy side --
 * Generator.uniform / integers / random streams that shape the terrain -- IS pinned against numpy
 * (tests/test_oracle_golden.py).
 */
#include "walker_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "b2lite.h"
#include "np_rng.h"

#define NB 5   /* 0 hull, 1 leg(-1), 2 lower(-1), 3 leg(+1), 4 lower(+1) */
#define NJ 4   /* 0 hip(-1), 1 knee(-1), 2 hip(+1), 3 knee(+1)  (self.joints order) */
#define NE 199 /* terrain edges i -> i+1 */

#define FPS 50
#define SCALE 30.0
#define MOTORS_TORQUE 80
#define SPEED_HIP 4
#define SPEED_KNEE 6
#define LIDAR_RANGE (160 / SCALE)
#define INITIAL_RANDOM 5
#define LEG_DOWN (-8 / SCALE)
#define LEG_W (8 / SCALE)
#define LEG_H (34 / SCALE)
#define VIEWPORT_W 600
#define VIEWPORT_H 400
#define TERRAIN_STEP (14 / SCALE)
#define TERRAIN_LENGTH 200
#define TERRAIN_HEIGHT (VIEWPORT_H / SCALE / 4)
#define TERRAIN_GRASS 10
#define TERRAIN_STARTPAD 20
#define FRICTION 2.5f

typedef struct {
    body_t b[NB];
    joint_t j[NJ];
    edge_t e[NE];
    contact_t ct[NB * NE];
    spoly_t sp[B2L_MAX_SPOLY];          /* hardcore obstacles, in creation (ascending x) order */
    int np;
    contact_t ctp[NB * B2L_MAX_SPOLY];
    float inv_dt0;
    int game_over, leg_contact[2]; /* legs[1] = body 2, legs[3] = body 4 */
    int has_prev_shaping;
    double prev_shaping;
    pcg64_t rng;
    int32_t elapsed;
    int stat_contacts, stat_pos_iters; /* of the last world step (workload statistics for DESIGN.md) */
    long long toi_calls, toi_events;   /* b2TimeOfImpact evaluations / TOI sub-steps since the env was created (sticky) */
    int toi_events_max;                /* most TOI sub-steps this env ever ran in ONE world step (sticky) */
    long long overflows;               /* touching pairs dropped because the scene's manifold table was full (sticky) */
} wworld_t;

struct orc_walker {
    int64_t n;
    int max_steps;
    int hardcore;
    wworld_t *w;
};

/* ContactDetector (bipedal_walker.py:80-98) */
/* manifold-table capacity of this scene: the CUDA scene's kMaxVC (gym_b200/csrc/walker.cuh); settable so that a test can
 * force overflows on both sides and check that they are handled identically */
static int g_walker_max_contacts = 10;
void orc_walker_set_max_contacts(int cap) { g_walker_max_contacts = cap; }
/* continuous collision (b2World::SolveTOI, b2lite_toi.h); on like Box2D's m_continuousPhysics, switchable so that a test
 * can show what it prevents */
static int g_walker_toi = 1;
void orc_walker_set_toi(int on) { g_walker_toi = on; }

static void walker_event(void *ctx, int body, int begin)
{
    wworld_t *W = (wworld_t *)ctx;
    if (begin) {
        if (body == 0) W->game_over = 1;
        if (body == 2) W->leg_contact[0] = 1;
        if (body == 4) W->leg_contact[1] = 1;
    } else {
        if (body == 2) W->leg_contact[0] = 0;
        if (body == 4) W->leg_contact[1] = 0;
    }
}

/* world.Step(1/50, 180, 60).  Island order from b2World::Solve's depth-first walk starting at the
 * newest body: lower(+1), leg(+1), hull, leg(-1), lower(-1); joints knee(+1), hip(+1), hip(-1), knee(-1). */
static void wworld_step(wworld_t *W)
{
    static const int order[NB] = {4, 3, 0, 1, 2}, jorder[NJ] = {3, 2, 0, 1};
    b2l_world S;
    S.nb = NB; S.nj = NJ; S.ne = NE;
    S.b = W->b; S.j = W->j; S.e = W->e; S.ct = W->ct;
    S.body_order = order; S.joint_order = jorder;
    S.np = W->np; S.np_cap = B2L_MAX_SPOLY; S.sp = W->sp; S.ctp = W->ctp;
    S.inv_dt0 = W->inv_dt0;
    S.gravity_y = -10.0f;
    S.event = walker_event; S.ctx = W;
    S.max_contacts = g_walker_max_contacts;
    S.toi = g_walker_toi; S.one_static_body = 0;
    b2l_step(&S, (float)(1.0 / FPS), 6 * 30, 2 * 30);
    W->inv_dt0 = S.inv_dt0;
    W->stat_contacts = S.stat_contacts; W->stat_pos_iters = S.stat_pos_iters;
    W->toi_calls += S.stat_toi_calls; W->toi_events += S.stat_toi_events;
    if (S.stat_toi_events > W->toi_events_max) W->toi_events_max = S.stat_toi_events;
    W->overflows += S.overflowed;
    /* step() rewrites every joint's motor each call, which wakes both bodies (b2RevoluteJoint::
     * SetMotorSpeed -> SetAwake(true)): an island that fell asleep is simply awake again next step */
    for (int i = 0; i < NB; i++) W->b[i].awake = 1;
}

/* b2EdgeShape::RayCast against terrain edge e (identity transform); returns 1 and *t on a hit */
static int edge_raycast(const edge_t *e, v2 p1, v2 p2, float maxFraction, float *t_out)
{
    v2 d = sub(p2, p1);
    v2 v1 = e->v1, v2_ = e->v2;
    v2 ed = sub(v2_, v1);
    v2 normal = V(ed.y, -ed.x);
    float len = sqrtf(normal.x * normal.x + normal.y * normal.y);
    if (len >= 1.1920929e-07f) { float inv = 1.0f / len; normal.x *= inv; normal.y *= inv; }
    float numerator = dot(normal, sub(v1, p1));
    float denominator = dot(normal, d);
    if (denominator == 0.0f) return 0;
    float t = numerator / denominator;
    if (t < 0.0f || maxFraction < t) return 0;
    v2 q = add(p1, scl(t, d));
    v2 r = sub(v2_, v1);
    float rr = dot(r, r);
    if (rr == 0.0f) return 0;
    float s = dot(sub(q, v1), r) / rr;
    if (s < 0.0f || 1.0f < s) return 0;
    *t_out = t;
    return 1;
}

static void walker_step_one(wworld_t *W, const float *action, int from_reset, float *obs, double *reward, int *terminated);

/* bipedal_walker.py:425-515 */
static void walker_reset_one(wworld_t *W, int hardcore, float *obs)
{
    pcg64_t rng = W->rng;
    float inv_dt0 = W->inv_dt0; /* self.world survives reset() */
    long long overflows = W->overflows, toi_calls = W->toi_calls, toi_events = W->toi_events;
    int toi_events_max = W->toi_events_max;
    memset(W, 0, sizeof *W);
    W->inv_dt0 = inv_dt0;
    W->overflows = overflows; W->toi_calls = toi_calls; W->toi_events = toi_events; W->toi_events_max = toi_events_max;
    /* _generate_terrain(hardcore) :277-402 */
    double terrain_x[TERRAIN_LENGTH], terrain_y[TERRAIN_LENGTH];
    {
        enum { GRASS = 0, STUMP, STAIRS, PIT, STATES };
        int state = GRASS;
        double velocity = 0.0, y = TERRAIN_HEIGHT, original_y = 0;
        int64_t counter = TERRAIN_STARTPAD, stair_steps = 0, stair_width = 0, stair_height = 0;
        int oneshot = 0;
        for (int i = 0; i < TERRAIN_LENGTH; i++) {
            double x = i * TERRAIN_STEP;
            terrain_x[i] = x;
            if (state == GRASS && !oneshot) {
                double d = TERRAIN_HEIGHT - y;
                double sgn = d > 0 ? 1.0 : (d < 0 ? -1.0 : 0.0);
                velocity = 0.8 * velocity + 0.01 * sgn;                              /* :296 */
                if (i > TERRAIN_STARTPAD) velocity += rng_uniform(&rng, -1, 1) / SCALE; /* :297-298 */
                y += velocity;
            } else if (state == PIT && oneshot) {                                     /* :301-323 */
                counter = rng_integers(&rng, 3, 5);
                /* poly = (x, y), (x+STEP, y), (x+STEP, y-4 STEP), (x, y-4 STEP), and the same shifted by STEP*counter */
                spoly_set_box(&W->sp[W->np++], (float)x, (float)(y - 4 * TERRAIN_STEP), (float)(x + TERRAIN_STEP), (float)y, FRICTION);
                spoly_set_box(&W->sp[W->np++], (float)(x + TERRAIN_STEP * counter), (float)(y - 4 * TERRAIN_STEP),
                              (float)(x + TERRAIN_STEP + TERRAIN_STEP * counter), (float)y, FRICTION);
                counter += 2;
                original_y = y;
            } else if (state == PIT && !oneshot) {                                    /* :325-328 */
                y = original_y;
                if (counter > 1) y -= 4 * TERRAIN_STEP;
            } else if (state == STUMP && oneshot) {                                   /* :330-341 */
                counter = rng_integers(&rng, 1, 3);
                spoly_set_box(&W->sp[W->np++], (float)x, (float)y, (float)(x + counter * TERRAIN_STEP),
                              (float)(y + counter * TERRAIN_STEP), FRICTION);
            } else if (state == STAIRS && oneshot) {                                  /* :343-371 */
                stair_height = pcg_double(&rng) > 0.5 ? +1 : -1;
                stair_width = rng_integers(&rng, 4, 5);
                stair_steps = rng_integers(&rng, 3, 5);
                original_y = y;
                for (int64_t st = 0; st < stair_steps; st++)
                    spoly_set_box(&W->sp[W->np++], (float)(x + (st * stair_width) * TERRAIN_STEP),
                                  (float)(y + (-1 + st * stair_height) * TERRAIN_STEP),
                                  (float)(x + ((1 + st) * stair_width) * TERRAIN_STEP),
                                  (float)(y + (st * stair_height) * TERRAIN_STEP), FRICTION);
                counter = stair_steps * stair_width;
            } else if (state == STAIRS && !oneshot) {                                 /* :373-376 */
                int64_t sq = stair_steps * stair_width - counter - stair_height;
                double nn = (double)sq / (double)stair_width;
                y = original_y + (nn * stair_height) * TERRAIN_STEP;
            }
            oneshot = 0;
            terrain_y[i] = y;
            counter -= 1;
            if (counter == 0) {
                counter = rng_integers(&rng, TERRAIN_GRASS / 2, TERRAIN_GRASS);       /* :379 */
                if (state == GRASS && hardcore) { state = (int)rng_integers(&rng, 1, STATES); oneshot = 1; }  /* :380-382 */
                else { state = GRASS; oneshot = 1; }                                   /* :383-385 */
            }
        }
        if (W->np > B2L_MAX_SPOLY) abort();  /* cannot happen: at most 39 */
    }
    for (int i = 0; i < NE; i++) {                                                     /* :387-396 */
        W->e[i].v1 = V((float)terrain_x[i], (float)terrain_y[i]);
        W->e[i].v2 = V((float)terrain_x[i + 1], (float)terrain_y[i + 1]);
        W->e[i].friction = FRICTION;
    }
    /* _generate_clouds :404-423: cosmetic, but it consumes the stream */
    for (int i = 0; i < TERRAIN_LENGTH / 20; i++) {
        (void)rng_uniform(&rng, 0, TERRAIN_LENGTH);
        for (int a = 0; a < 5; a++) { (void)rng_uniform(&rng, 0, 5 * TERRAIN_STEP); (void)rng_uniform(&rng, 0, 5 * TERRAIN_STEP); }
    }
    const double init_x = TERRAIN_STEP * TERRAIN_STARTPAD / 2, init_y = TERRAIN_HEIGHT + 2 * LEG_H; /* :442-443 */
    /* hull :444-452; vertex order as b2PolygonShape::Set's gift wrapping leaves it */
    static const double HP[5][2] = {{34, -8}, {34, 1}, {6, 9}, {-30, 9}, {-30, -8}};
    v2 hull[5];
    for (int i = 0; i < 5; i++) hull[i] = V((float)(HP[i][0] / SCALE), (float)(HP[i][1] / SCALE));
    poly_set(&W->b[0], hull, 5);
    poly_mass(&W->b[0], 5.0f);
    W->b[0].friction = 0.1f;
    body_place(&W->b[0], V((float)init_x, (float)init_y), 0.0f);
    double fx = rng_uniform(&rng, -INITIAL_RANDOM, INITIAL_RANDOM);
    W->b[0].force = add(W->b[0].force, V((float)fx, 0.0f));
    for (int li = 0; li < 2; li++) {                                                   /* :456-500 */
        int i = li == 0 ? -1 : +1;
        body_t *leg = &W->b[1 + 2 * li], *lower = &W->b[2 + 2 * li];
        {
            float hx = (float)(LEG_W / 2), hy = (float)(LEG_H / 2);
            v2 box[4] = {V(-hx, -hy), V(hx, -hy), V(hx, hy), V(-hx, hy)};
            poly_set(leg, box, 4);
            leg->normals[0] = V(0.0f, -1.0f); leg->normals[1] = V(1.0f, 0.0f); leg->normals[2] = V(0.0f, 1.0f); leg->normals[3] = V(-1.0f, 0.0f);
            leg->centroid = V(0.0f, 0.0f);
            poly_mass(leg, 1.0f);
            leg->friction = 0.2f;
            body_place(leg, V((float)init_x, (float)(init_y - LEG_H / 2 - LEG_DOWN)), (float)(i * 0.05));
        }
        {
            float hx = (float)(0.8 * LEG_W / 2), hy = (float)(LEG_H / 2);
            v2 box[4] = {V(-hx, -hy), V(hx, -hy), V(hx, hy), V(-hx, hy)};
            poly_set(lower, box, 4);
            lower->normals[0] = V(0.0f, -1.0f); lower->normals[1] = V(1.0f, 0.0f); lower->normals[2] = V(0.0f, 1.0f); lower->normals[3] = V(-1.0f, 0.0f);
            lower->centroid = V(0.0f, 0.0f);
            poly_mass(lower, 1.0f);
            lower->friction = 0.2f;
            body_place(lower, V((float)init_x, (float)(init_y - LEG_H * 3 / 2 - LEG_DOWN)), (float)(i * 0.05));
        }
        joint_t *hip = &W->j[2 * li], *knee = &W->j[2 * li + 1];
        hip->bodyA = 0; hip->bodyB = 1 + 2 * li;
        hip->localAnchorA = V(0.0f, (float)LEG_DOWN); hip->localAnchorB = V(0.0f, (float)(LEG_H / 2));
        hip->maxMotorTorque = (float)MOTORS_TORQUE; hip->motorSpeed = (float)i;
        hip->lower = -0.8f; hip->upper = 1.1f; hip->referenceAngle = 0.0f;
        knee->bodyA = 1 + 2 * li; knee->bodyB = 2 + 2 * li;
        knee->localAnchorA = V(0.0f, (float)(-LEG_H / 2)); knee->localAnchorB = V(0.0f, (float)(LEG_H / 2));
        knee->maxMotorTorque = (float)MOTORS_TORQUE; knee->motorSpeed = 1.0f;
        knee->lower = -1.6f; knee->upper = -0.1f; knee->referenceAngle = 0.0f;
    }
    W->rng = rng;
    W->elapsed = 0;
    const float zero[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    double r;
    int t;
    walker_step_one(W, zero, 1, obs, &r, &t);                                          /* :515 */
}

/* np.sign / np.clip(np.abs(a), 0, 1) on a float32 action component */
static inline float sgnf(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
static inline float clip01_abs(float a) { float x = fabsf(a); x = x < 0.0f ? 0.0f : x; return x > 1.0f ? 1.0f : x; }

/* bipedal_walker.py:517-606 */
static void walker_step_one(wworld_t *W, const float *action, int from_reset, float *obs, double *reward, int *terminated)
{
    static const float speed[NJ] = {SPEED_HIP, SPEED_KNEE, SPEED_HIP, SPEED_KNEE};
    for (int k = 0; k < NJ; k++) {                                                     /* :528-543 */
        W->j[k].motorSpeed = speed[k] * sgnf(action[k]);
        W->j[k].maxMotorTorque = (float)MOTORS_TORQUE * clip01_abs(action[k]);
    }
    wworld_step(W);                                                                    /* :545 */
    body_t *H = &W->b[0];
    const double posx = (double)H->xf.p.x, posy = (double)H->xf.p.y;
    float lidar[10];
    for (int i = 0; i < 10; i++) {                                                     /* :550-557 */
        double p2x = posx + sin(1.5 * i / 10.0) * LIDAR_RANGE, p2y = posy - cos(1.5 * i / 10.0) * LIDAR_RANGE;
        v2 p1 = V((float)posx, (float)posy), p2 = V((float)p2x, (float)p2y);
        float frac = 1.0f, maxFraction = 1.0f;
        for (int e = 0; e < NE; e++) {
            float t;
            if (edge_raycast(&W->e[e], p1, p2, maxFraction, &t)) { frac = t; maxFraction = t; }
        }
        for (int q = 0; q < W->np; q++) {  /* LidarCallback accepts every fixture with categoryBits & 1 (:504-510) */
            float t;
            if (spoly_raycast(&W->sp[q], p1, p2, maxFraction, &t)) { frac = t; maxFraction = t; }
        }
        lidar[i] = frac;
    }
    double st[24];
    st[0] = (double)H->a;                                                              /* :559-577 */
    st[1] = 2.0 * (double)H->w / FPS;
    st[2] = 0.3 * (double)H->v.x * (VIEWPORT_W / SCALE) / FPS;
    st[3] = 0.3 * (double)H->v.y * (VIEWPORT_H / SCALE) / FPS;
    for (int li = 0; li < 2; li++) {
        const joint_t *hip = &W->j[2 * li], *knee = &W->j[2 * li + 1];
        float hip_angle = W->b[hip->bodyB].a - W->b[hip->bodyA].a - hip->referenceAngle;     /* GetJointAngle */
        float hip_speed = W->b[hip->bodyB].w - W->b[hip->bodyA].w;                           /* GetJointSpeed */
        float knee_angle = W->b[knee->bodyB].a - W->b[knee->bodyA].a - knee->referenceAngle;
        float knee_speed = W->b[knee->bodyB].w - W->b[knee->bodyA].w;
        st[4 + 5 * li] = (double)hip_angle;
        st[5 + 5 * li] = (double)hip_speed / SPEED_HIP;
        st[6 + 5 * li] = (double)knee_angle + 1.0;
        st[7 + 5 * li] = (double)knee_speed / SPEED_KNEE;
        st[8 + 5 * li] = W->leg_contact[li] ? 1.0 : 0.0;
    }
    for (int i = 0; i < 10; i++) st[14 + i] = (double)lidar[i];
    double shaping = 130 * posx / SCALE;                                               /* :582-587 */
    shaping -= 5.0 * fabs(st[0]);
    double r = 0;
    if (W->has_prev_shaping) r = shaping - W->prev_shaping;                            /* :589-592 */
    W->prev_shaping = shaping;
    W->has_prev_shaping = 1;
    /* :594-596 -- under numpy >= 2 `python_float - np.float32` is float32: the torque costs are
     * subtracted in float32 (0.00035 * MOTORS_TORQUE is a weak Python scalar) */
    double rew = r;
    if (!from_reset) {
        float r32 = (float)r;
        for (int k = 0; k < NJ; k++) r32 = r32 - (float)(0.00035 * MOTORS_TORQUE) * clip01_abs(action[k]);
        rew = (double)r32;
    }
    int term = 0;
    if (W->game_over || posx < 0) { rew = -100; term = 1; }                            /* :598-601 */
    if (posx > (TERRAIN_LENGTH - TERRAIN_GRASS) * TERRAIN_STEP) term = 1;              /* :602-603 */
    for (int k = 0; k < 24; k++) obs[k] = (float)st[k];                                /* :606 */
    *reward = rew;
    *terminated = term;
}

/* test hook: the action-dependent arithmetic of step() alone (:528-543,594-596) -- motor speed / torque of the four
 * joints as Box2D receives them and the reward after the torque costs, given the shaping delta */
void orc_walker_action_flow(double shaping_delta, const float action[4], float motor_speed[4], float max_torque[4],
                            double *reward)
{
    static const float speed[NJ] = {SPEED_HIP, SPEED_KNEE, SPEED_HIP, SPEED_KNEE};
    for (int k = 0; k < NJ; k++) {
        motor_speed[k] = speed[k] * sgnf(action[k]);
        max_torque[k] = (float)MOTORS_TORQUE * clip01_abs(action[k]);
    }
    float r32 = (float)shaping_delta;
    for (int k = 0; k < NJ; k++) r32 = r32 - (float)(0.00035 * MOTORS_TORQUE) * clip01_abs(action[k]);
    *reward = (double)r32;
}

/* ---------------------------------------------------------------- vector API */
orc_walker *orc_walker_create(int64_t n, int max_episode_steps) { return orc_walker_create_ex(n, max_episode_steps, 0); }

orc_walker *orc_walker_create_ex(int64_t n, int max_episode_steps, int hardcore)
{
    if (n <= 0) return NULL;
    orc_walker *v = (orc_walker *)calloc(1, sizeof *v);
    v->n = n;
    v->max_steps = max_episode_steps;
    v->hardcore = hardcore;
    v->w = (wworld_t *)calloc((size_t)n, sizeof(wworld_t));
    return v;
}
void orc_walker_destroy(orc_walker *v) { if (v) { free(v->w); free(v); } }

void orc_walker_seed_range(orc_walker *v, const uint32_t base[4], int64_t first)
{
    u128 b = 0;
    for (int k = 3; k >= 0; k--) b = (b << 32) | base[k];
    for (int64_t i = 0; i < v->n; i++) {
        u128 s = b + (u128)(uint64_t)(first + i);
        uint32_t ent[4] = {(uint32_t)s, (uint32_t)(s >> 32), (uint32_t)(s >> 64), (uint32_t)(s >> 96)};
        pcg_seed_from_words(&v->w[i].rng, ent);
    }
}

void orc_walker_reset(orc_walker *v, float *obs)
{
    for (int64_t i = 0; i < v->n; i++) walker_reset_one(&v->w[i], v->hardcore, obs + 24 * i);
}

typedef struct {
    orc_walker *v; const float *actions; float *obs; double *reward; uint8_t *terminated, *truncated; float *final_obs;
    int64_t lo, hi;
} walker_job;

static void *walker_range(void *arg)
{
    walker_job *j = (walker_job *)arg;
    orc_walker *v = j->v;
    const float *actions = j->actions;
    float *obs = j->obs, *final_obs = j->final_obs;
    double *reward = j->reward;
    uint8_t *terminated = j->terminated, *truncated = j->truncated;
    for (int64_t i = j->lo; i < j->hi; i++) {
        wworld_t *W = &v->w[i];
        float o[24];
        double r;
        int term;
        walker_step_one(W, actions + 4 * i, 0, o, &r, &term);
        W->elapsed += 1;
        int trunc = v->max_steps > 0 && W->elapsed >= v->max_steps;
        reward[i] = r;
        terminated[i] = (uint8_t)term;
        truncated[i] = (uint8_t)trunc;
        if (term || trunc) {
            if (final_obs) memcpy(final_obs + 24 * i, o, sizeof o);
            walker_reset_one(W, v->hardcore, o);
        }
        memcpy(obs + 24 * i, o, sizeof o);
    }
    return NULL;
}

/* the envs are independent: a range of them per host thread (bench.py's cpu_baseline / reference arm) */
void orc_walker_step_mt(orc_walker *v, const float *actions, float *obs, double *reward, uint8_t *terminated,
                        uint8_t *truncated, float *final_obs, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > v->n) nthreads = (int)v->n;
    walker_job jobs[256];
    pthread_t tid[256];
    const int64_t chunk = (v->n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        walker_job *j = &jobs[t];
        j->v = v; j->actions = actions; j->obs = obs; j->reward = reward; j->terminated = terminated;
        j->truncated = truncated; j->final_obs = final_obs;
        j->lo = t * chunk;
        j->hi = (j->lo + chunk < v->n) ? j->lo + chunk : v->n;
        if (j->lo > j->hi) j->lo = j->hi;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&tid[t], NULL, walker_range, &jobs[t]);
    walker_range(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(tid[t], NULL);
}

void orc_walker_step(orc_walker *v, const float *actions, float *obs, double *reward, uint8_t *terminated,
                     uint8_t *truncated, float *final_obs)
{
    orc_walker_step_mt(v, actions, obs, reward, terminated, truncated, final_obs, 1);
}

/* envs in which a touching pair was ever dropped because the manifold table was full */
int64_t orc_walker_overflows(const orc_walker *v)
{
    int64_t c = 0;
    for (int64_t i = 0; i < v->n; i++) c += v->w[i].overflows != 0;
    return c;
}

/* {b2TimeOfImpact evaluations, TOI sub-steps} summed over all envs since creation, {most sub-steps in one world step} */
void orc_walker_toi_stats(const orc_walker *v, int64_t out[3])
{
    out[0] = 0; out[1] = 0; out[2] = 0;
    for (int64_t i = 0; i < v->n; i++) {
        out[0] += v->w[i].toi_calls; out[1] += v->w[i].toi_events;
        if (v->w[i].toi_events_max > out[2]) out[2] = v->w[i].toi_events_max;
    }
}

void orc_walker_get_stats(const orc_walker *v, int32_t *out)
{
    for (int64_t i = 0; i < v->n; i++) { out[2 * i] = v->w[i].stat_contacts; out[2 * i + 1] = v->w[i].stat_pos_iters; }
}

void orc_walker_get_terrain(const orc_walker *v, int64_t i, float *y200)
{
    const wworld_t *W = &v->w[i];
    for (int k = 0; k < NE; k++) y200[k] = W->e[k].v1.y;
    y200[NE] = W->e[NE - 1].v2.y;
}

/* hardcore obstacles of env i: out[k] = {x0, ylo, x1, yhi}; returns their number */
int orc_walker_get_polys(const orc_walker *v, int64_t i, float *out)
{
    const wworld_t *W = &v->w[i];
    for (int k = 0; k < W->np; k++) {
        out[4 * k + 0] = W->sp[k].verts[2].x; out[4 * k + 1] = W->sp[k].verts[0].y;
        out[4 * k + 2] = W->sp[k].verts[0].x; out[4 * k + 3] = W->sp[k].verts[2].y;
    }
    return W->np;
}

void orc_walker_get_bodies(const orc_walker *v, int64_t i, float out[30], int32_t flags[4])
{
    const wworld_t *W = &v->w[i];
    for (int b = 0; b < NB; b++) {
        out[6 * b + 0] = W->b[b].c.x; out[6 * b + 1] = W->b[b].c.y; out[6 * b + 2] = W->b[b].a;
        out[6 * b + 3] = W->b[b].v.x; out[6 * b + 4] = W->b[b].v.y; out[6 * b + 5] = W->b[b].w;
    }
    flags[0] = W->game_over; flags[1] = W->leg_contact[0]; flags[2] = W->leg_contact[1]; flags[3] = 0;
    for (int k = 0; k < NB * NE; k++) flags[3] += W->ct[k].touching;
    for (int k = 0; k < NB * B2L_MAX_SPOLY; k++) flags[3] += W->ctp[k].touching;
}

/* Generator KATs for tests: draws[k] = op(k): 0 uniform(-1,1), 1 integers(5,10), 2 random() */
void orc_rng_sequence(const uint32_t ent[4], const int32_t *ops, int64_t n, double *out)
{
    pcg64_t g;
    pcg_seed_from_words(&g, ent);
    for (int64_t k = 0; k < n; k++) {
        if (ops[k] == 0) out[k] = rng_uniform(&g, -1, 1);
        else if (ops[k] == 1) out[k] = (double)rng_integers(&g, 5, 10);
        else if (ops[k] == 3) out[k] = (double)rng_integers(&g, 1, 5);
        else out[k] = pcg_double(&g);
    }
}
