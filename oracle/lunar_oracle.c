/*
 * lunar_oracle.c -- CPU ORACLE for LunarLander-v2 (test infrastructure, NOT product code).
 *
 * Restates gym/envs/box2d/lunar_lander.py (reset :308-420, step :444-600, ContactDetector
 * :54-72) on top of a from-scratch restatement of the part of Box2D 2.3.x that the env
 * exercises.  box2d-py == 2.3.5 (setup.py:15) is a third-party dependency that is neither
 * vendored in /root/reference nor installable here, so its algorithm is restated from its
 * published design (Erin Catto's sequential-impulse solver as shipped in Box2D 2.3):
 *   b2World::Step(dt, 180, 60): Collide (edge-vs-polygon manifolds, begin/end events, warm-start
 *   impulse matching by feature id) -> island solve (integrate velocities, revolute joints with
 *   limit + motor, 2-point block contact solver with friction, 180 velocity iterations, position
 *   integration with translation clamps, <= 60 position iterations with Baumgarte 0.2 / slop
 *   0.005) -> sleeping (0.5 s under 0.01 m/s, 2 deg/s).
 *
 * PARITY UNPINNED: no Box2D build and no golden vectors exist for this path (the only reference
 * test at this boundary is the behavioural tests/envs/test_env_implementation.py:13-17, heuristic
 * return > 100 at seed 1).  Known, deliberate deviations from Box2D: continuous collision (TOI
 * sub-stepping) is not modelled; broad-phase pair bookkeeping is replaced by testing every
 * (body, edge) pair every step (same manifolds); cosmetic particles are not simulated (they
 * collide only with the ground: categoryBits 0x0100 vs the lander/leg maskBits 0x001); contacts
 * inside an island are ordered (leg+1, lander, leg-1) x (edge 10..0); sin/cos of body angles use
 * the polynomial below instead of the platform libm (Box2D's own results are libm-dependent).
 *
 * All arithmetic inside the physics is IEEE float32 with separate rounding of every operation
 * (build with -ffp-contract=off) so that the CUDA kernel can be compared bit for bit.
 */
#include "lunar_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "b2lite.h"
#include "np_rng.h"

#define NB 3    /* dynamic bodies: 0 lander, 1 leg (i=-1), 2 leg (i=+1) */
#define NE 11   /* ground edges: 0 base, 1..10 terrain */

typedef struct {
    body_t b[NB];
    edge_t e[NE];
    contact_t ct[NB * NE];
    joint_t j[2];
    float inv_dt0;
    int game_over, leg_contact[2];
    int has_prev_shaping;
    double prev_shaping, helipad_y;
    pcg64_t rng;
    int32_t elapsed;
    int stat_contacts, stat_pos_iters; /* of the last world step (workload statistics for DESIGN.md) */
    long long toi_calls, toi_events;   /* b2TimeOfImpact evaluations / TOI sub-steps since the env was created (sticky) */
    int toi_events_max;                /* most TOI sub-steps this env ever ran in ONE world step (sticky) */
    long long overflows;               /* touching pairs dropped because the scene's manifold table was full (sticky) */
    int32_t wind_idx, torque_idx; /* lunar_lander.py:234-235: drawn once per env object, never reset */
} world_t;

typedef struct {
    int continuous, enable_wind;              /* lunar_lander.py:194-199 */
    double gravity, wind_power, turbulence_power;
} opts_t;

struct orc_lunar {
    int64_t n;
    int max_steps;
    opts_t o;
    world_t *w;
};

/* ContactDetector (lunar_lander.py:54-72) */
/* manifold-table capacity of this scene: the CUDA scene's kMaxVC (gym_b200/csrc/lunar.cuh); settable so that a test can
 * force overflows on both sides and check that they are handled identically */
static int g_lunar_max_contacts = 8;
void orc_lunar_set_max_contacts(int cap) { g_lunar_max_contacts = cap; }
/* continuous collision (b2World::SolveTOI, b2lite_toi.h); on like Box2D's m_continuousPhysics, switchable so that a test
 * can show what it prevents */
static int g_lunar_toi = 1;
void orc_lunar_set_toi(int on) { g_lunar_toi = on; }

static void lunar_event(void *ctx, int body, int begin)
{
    world_t *W = (world_t *)ctx;
    if (begin) { if (body == 0) W->game_over = 1; else W->leg_contact[body - 1] = 1; }
    else if (body > 0) W->leg_contact[body - 1] = 0;
}

/* world.Step(1/50, 180, 60): island order {leg+1, lander, leg-1}, joint of leg+1 first */
static int world_step(world_t *W, float gravity_y, float dt, int velIters, int posIters)
{
    static const int order[NB] = {2, 0, 1}, jorder[2] = {1, 0};
    b2l_world S;
    S.nb = NB; S.nj = 2; S.ne = NE;
    S.b = W->b; S.j = W->j; S.e = W->e; S.ct = W->ct;
    S.body_order = order; S.joint_order = jorder;
    S.np = 0; S.np_cap = 0; S.sp = NULL; S.ctp = NULL;
    S.inv_dt0 = W->inv_dt0;
    S.gravity_y = gravity_y;
    S.event = lunar_event; S.ctx = W;
    S.max_contacts = g_lunar_max_contacts;
    S.toi = g_lunar_toi; S.one_static_body = 1;
    b2l_step(&S, dt, velIters, posIters);
    W->inv_dt0 = S.inv_dt0;
    W->stat_contacts = S.stat_contacts; W->stat_pos_iters = S.stat_pos_iters;
    W->toi_calls += S.stat_toi_calls; W->toi_events += S.stat_toi_events;
    if (S.stat_toi_events > W->toi_events_max) W->toi_events_max = S.stat_toi_events;
    W->overflows += S.overflowed;
    return S.awake;
}

/* ---------------------------------------------------------------- the environment */
#define FPS 50
#define SCALE 30.0
#define VIEWPORT_W 600
#define VIEWPORT_H 400
#define MAIN_ENGINE_POWER 13.0
#define SIDE_ENGINE_POWER 0.6
#define INITIAL_RANDOM 1000.0
#define LEG_AWAY 20
#define LEG_DOWN 18
#define LEG_W 2
#define LEG_H 8
#define LEG_SPRING_TORQUE 40
#define SIDE_ENGINE_HEIGHT 14.0
#define SIDE_ENGINE_AWAY 12.0

static void lunar_step_one(world_t *W, const opts_t *O, int action, const float *caction, float *obs, double *reward,
                           int *terminated);

/* lunar_lander.py:308-420 */
static void lunar_reset_one(world_t *W, const opts_t *O, float *obs)
{
    pcg64_t rng = W->rng;
    float inv_dt0 = W->inv_dt0; /* the b2World object survives reset() */
    int32_t wind_idx = W->wind_idx, torque_idx = W->torque_idx;
    long long overflows = W->overflows, toi_calls = W->toi_calls, toi_events = W->toi_events;
    int toi_events_max = W->toi_events_max;
    memset(W, 0, sizeof *W);
    W->inv_dt0 = inv_dt0;
    W->overflows = overflows; W->toi_calls = toi_calls; W->toi_events = toi_events; W->toi_events_max = toi_events_max;
    W->wind_idx = wind_idx; W->torque_idx = torque_idx;
    const double Wd = VIEWPORT_W / SCALE, Hd = VIEWPORT_H / SCALE;
    enum { CHUNKS = 11 };
    double height[CHUNKS + 1], chunk_x[CHUNKS], smooth_y[CHUNKS];
    for (int i = 0; i < CHUNKS + 1; i++) height[i] = rng_uniform(&rng, 0, Hd / 2);          /* :326 */
    for (int i = 0; i < CHUNKS; i++) chunk_x[i] = Wd / (CHUNKS - 1) * i;                    /* :327 */
    W->helipad_y = Hd / 4;                                                                 /* :330 */
    for (int k = -2; k <= 2; k++) height[CHUNKS / 2 + k] = W->helipad_y;                    /* :331-335 */
    for (int i = 0; i < CHUNKS; i++)                                                        /* :336-339 */
        smooth_y[i] = 0.33 * (height[(i - 1 + CHUNKS + 1) % (CHUNKS + 1)] + height[i + 0] + height[i + 1]);
    W->e[0].v1 = V(0.0f, 0.0f); W->e[0].v2 = V((float)Wd, 0.0f); W->e[0].friction = 0.2f;  /* :341-343 */
    for (int i = 0; i < CHUNKS - 1; i++) {                                                  /* :345-349 */
        W->e[i + 1].v1 = V((float)chunk_x[i], (float)smooth_y[i]);
        W->e[i + 1].v2 = V((float)chunk_x[i + 1], (float)smooth_y[i + 1]);
        W->e[i + 1].friction = 0.1f;
    }
    const double initial_y = VIEWPORT_H / SCALE;
    /* lander :354-368; hull order as b2PolygonShape::Set produces it */
    static const double LP[6][2] = {{17, -10}, {17, 0}, {14, 17}, {-14, 17}, {-17, 0}, {-17, -10}};
    v2 hull[6];
    for (int i = 0; i < 6; i++) hull[i] = V((float)(LP[i][0] / SCALE), (float)(LP[i][1] / SCALE));
    poly_set(&W->b[0], hull, 6);
    poly_mass(&W->b[0], 5.0f);
    W->b[0].friction = 0.1f;
    body_place(&W->b[0], V((float)(VIEWPORT_W / SCALE / 2), (float)initial_y), 0.0f);
    double fx = rng_uniform(&rng, -INITIAL_RANDOM, INITIAL_RANDOM);                        /* :371-377 */
    double fy = rng_uniform(&rng, -INITIAL_RANDOM, INITIAL_RANDOM);
    W->b[0].force = add(W->b[0].force, V((float)fx, (float)fy));
    /* legs :379-414 */
    for (int li = 0; li < 2; li++) {
        int i = li == 0 ? -1 : +1;
        body_t *leg = &W->b[1 + li];
        float hx = (float)(LEG_W / SCALE), hy = (float)(LEG_H / SCALE);
        v2 box[4] = {V(-hx, -hy), V(hx, -hy), V(hx, hy), V(-hx, hy)};
        poly_set(leg, box, 4);
        /* SetAsBox writes exact normals and a zero centroid */
        leg->normals[0] = V(0.0f, -1.0f); leg->normals[1] = V(1.0f, 0.0f); leg->normals[2] = V(0.0f, 1.0f); leg->normals[3] = V(-1.0f, 0.0f);
        leg->centroid = V(0.0f, 0.0f);
        poly_mass(leg, 1.0f);
        leg->friction = 0.2f;
        body_place(leg, V((float)(VIEWPORT_W / SCALE / 2 - i * LEG_AWAY / SCALE), (float)initial_y), (float)(i * 0.05));
        joint_t *j = &W->j[li];
        j->bodyA = 0;
        j->bodyB = 1 + li;
        j->localAnchorA = V(0.0f, 0.0f);
        j->localAnchorB = V((float)(i * LEG_AWAY / SCALE), (float)(LEG_DOWN / SCALE));
        j->maxMotorTorque = (float)LEG_SPRING_TORQUE;
        j->motorSpeed = (float)(+0.3 * i);
        if (i == -1) { j->lower = (float)(+0.9 - 0.5); j->upper = (float)(+0.9); }
        else { j->lower = (float)(-0.9); j->upper = (float)(-0.9 + 0.5); }
        j->referenceAngle = 0.0f;
    }
    W->rng = rng;
    W->elapsed = 0;
    double r;
    int t;
    const float zero[2] = {0.0f, 0.0f};
    lunar_step_one(W, O, 0, zero, obs, &r, &t);                                            /* :420 */
}

/* The two engines of lunar_lander.py:486-554: impulse on the lander and its application point, both as the
 * float32 pairs Box2D receives.  The discrete branch is Python-float (f64) arithmetic rounded once on entry
 * to Box2D; in the continuous branch m_power / s_power / direction are numpy float32 scalars and, under
 * NEP 50 (numpy >= 2), every Python float they meet is first rounded to float32 -- the roundings below
 * are pinned against numpy itself by tests/test_oracle_box2d.py. */
typedef struct {
    int main_on, side_on;
    v2 main_imp, main_pt, side_imp, side_pt;
    double main_cost, side_cost;   /* m_power * 0.30, s_power * 0.03 as added to the reward (:585-588) */
} engines_t;

static void lunar_engines(engines_t *E, int continuous, int action, const float *ca, double ang, double posx, double posy,
                          double disp0, double disp1)
{
    double tip0 = sin(ang), tip1 = cos(ang);                                               /* :487 */
    double side0 = -tip1, side1 = tip0;
    memset(E, 0, sizeof *E);
    float a0 = 0.0f, a1 = 0.0f;
    if (continuous) {                                                                      /* :480 */
        a0 = ca[0] < -1.0f ? -1.0f : (ca[0] > 1.0f ? 1.0f : ca[0]);
        a1 = ca[1] < -1.0f ? -1.0f : (ca[1] > 1.0f ? 1.0f : ca[1]);
    }
    if (continuous ? (a0 > 0.0f) : (action == 2)) {                                        /* :491-520 */
        double ox = tip0 * (4 / SCALE + 2 * disp0) + side0 * disp1;
        double oy = -tip1 * (4 / SCALE + 2 * disp0) - side1 * disp1;
        E->main_on = 1;
        E->main_pt = V((float)(posx + ox), (float)(posy + oy));
        if (continuous) {
            float c = a0 < 0.0f ? 0.0f : (a0 > 1.0f ? 1.0f : a0);
            float m_power = (c + 1.0f) * 0.5f;                                             /* :497, float32 */
            E->main_imp = V((float)(-ox * MAIN_ENGINE_POWER) * m_power, (float)(-oy * MAIN_ENGINE_POWER) * m_power);
            E->main_cost = (double)(m_power * 0.30f);
        } else {
            E->main_imp = V((float)(-ox * MAIN_ENGINE_POWER * 1.0), (float)(-oy * MAIN_ENGINE_POWER * 1.0));
            E->main_cost = 1.0 * 0.30;
        }
    }
    if (continuous ? (fabsf(a1) > 0.5f) : (action == 1 || action == 3)) {                  /* :522-554 */
        E->side_on = 1;
        if (continuous) {
            float direction = a1 > 0.0f ? 1.0f : -1.0f;                                    /* np.sign, |a1| > 0.5 */
            float ab = fabsf(a1);
            float s_power = ab < 0.5f ? 0.5f : (ab > 1.0f ? 1.0f : ab);
            float t = direction * 12.0f / 30.0f;
            float u = (float)(3 * disp1) + t;
            float ox = (float)(tip0 * disp0) + (float)side0 * u;
            float oy = (float)(-tip1 * disp0) - (float)side1 * u;
            E->side_pt = V(((float)posx + ox) - (float)(tip0 * 17 / SCALE), ((float)posy + oy) + (float)(tip1 * SIDE_ENGINE_HEIGHT / SCALE));
            E->side_imp = V((-ox * 0.6f) * s_power, (-oy * 0.6f) * s_power);
            E->side_cost = (double)(s_power * 0.03f);
        } else {
            double direction = action - 2;
            double ox = tip0 * disp0 + side0 * (3 * disp1 + direction * SIDE_ENGINE_AWAY / SCALE);
            double oy = -tip1 * disp0 - side1 * (3 * disp1 + direction * SIDE_ENGINE_AWAY / SCALE);
            E->side_pt = V((float)(posx + ox - tip0 * 17 / SCALE), (float)(posy + oy + tip1 * SIDE_ENGINE_HEIGHT / SCALE));
            E->side_imp = V((float)(-ox * SIDE_ENGINE_POWER * 1.0), (float)(-oy * SIDE_ENGINE_POWER * 1.0));
            E->side_cost = 1.0 * 0.03;
        }
    }
}

/* test hook: the engine arithmetic alone (out = main imp.x, imp.y, pt.x, pt.y, side imp.x, imp.y, pt.x, pt.y) */
void orc_lunar_engines(int continuous, int action, const float *caction, double ang, double posx, double posy,
                       double disp0, double disp1, float out[8], double cost[2], int32_t on[2])
{
    engines_t E;
    lunar_engines(&E, continuous, action, caction, ang, posx, posy, disp0, disp1);
    out[0] = E.main_imp.x; out[1] = E.main_imp.y; out[2] = E.main_pt.x; out[3] = E.main_pt.y;
    out[4] = E.side_imp.x; out[5] = E.side_imp.y; out[6] = E.side_pt.x; out[7] = E.side_pt.y;
    cost[0] = E.main_cost; cost[1] = E.side_cost;
    on[0] = E.main_on; on[1] = E.side_on;
}

/* lunar_lander.py:444-600 */
static void lunar_step_one(world_t *W, const opts_t *O, int action, const float *caction, float *obs, double *reward,
                           int *terminated)
{
    body_t *L = &W->b[0];
    if (O->enable_wind && !(W->leg_contact[0] || W->leg_contact[1])) {                     /* :449-477 */
        double wind_mag = tanh(sin(0.02 * W->wind_idx) + sin(M_PI * 0.01 * W->wind_idx)) * O->wind_power;
        W->wind_idx += 1;
        L->force = add(L->force, V((float)wind_mag, 0.0f));                                /* ApplyForceToCenter */
        double torque_mag = tanh(sin(0.02 * W->torque_idx) + sin(M_PI * 0.01 * W->torque_idx)) * O->turbulence_power;
        W->torque_idx += 1;
        L->torque += (float)torque_mag;                                                    /* ApplyTorque */
    }
    double disp0 = rng_uniform(&W->rng, -1.0, +1.0) / SCALE;                               /* :489 */
    double disp1 = rng_uniform(&W->rng, -1.0, +1.0) / SCALE;
    engines_t E;
    lunar_engines(&E, O->continuous, action, caction, (double)L->a, (double)L->xf.p.x, (double)L->xf.p.y, disp0, disp1);
    if (E.main_on) { /* b2Body::ApplyLinearImpulse(impulse, point, wake=True) */
        L->v = add(L->v, scl(L->invMass, E.main_imp));
        L->w += L->invI * crs(sub(E.main_pt, L->c), E.main_imp);
    }
    if (E.side_on) {
        L->v = add(L->v, scl(L->invMass, E.side_imp));
        L->w += L->invI * crs(sub(E.side_pt, L->c), E.side_imp);
    }
    int awake = world_step(W, (float)O->gravity, (float)(1.0 / FPS), 6 * 30, 2 * 30);                         /* :556 */
    double st[8];
    st[0] = ((double)L->xf.p.x - VIEWPORT_W / SCALE / 2) / (VIEWPORT_W / SCALE / 2);       /* :560-569 */
    st[1] = ((double)L->xf.p.y - (W->helipad_y + LEG_DOWN / SCALE)) / (VIEWPORT_H / SCALE / 2);
    st[2] = (double)L->v.x * (VIEWPORT_W / SCALE / 2) / FPS;
    st[3] = (double)L->v.y * (VIEWPORT_H / SCALE / 2) / FPS;
    st[4] = (double)L->a;
    st[5] = 20.0 * (double)L->w / FPS;
    st[6] = W->leg_contact[0] ? 1.0 : 0.0;
    st[7] = W->leg_contact[1] ? 1.0 : 0.0;
    double r = 0;
    double shaping = -100 * sqrt(st[0] * st[0] + st[1] * st[1]) - 100 * sqrt(st[2] * st[2] + st[3] * st[3])
                     - 100 * fabs(st[4]) + 10 * st[6] + 10 * st[7];                        /* :572-578 */
    if (W->has_prev_shaping) r = shaping - W->prev_shaping;                                /* :581-583 */
    W->prev_shaping = shaping;
    W->has_prev_shaping = 1;
    r -= E.main_cost;                                                                      /* :585-588 */
    r -= E.side_cost;
    int term = 0;
    if (W->game_over || fabs(st[0]) >= 1.0) { term = 1; r = -100; }                        /* :590-593 */
    if (!awake) { term = 1; r = +100; }                                                 /* :594-596 */
    for (int k = 0; k < 8; k++) obs[k] = (float)st[k];                                     /* :600 */
    *reward = r;
    *terminated = term;
}

/* ---------------------------------------------------------------- vector API */
orc_lunar *orc_lunar_create_ex(int64_t n, int max_episode_steps, int continuous, int enable_wind, double gravity,
                               double wind_power, double turbulence_power)
{
    if (n <= 0) return NULL;
    orc_lunar *v = (orc_lunar *)calloc(1, sizeof *v);
    v->n = n;
    v->max_steps = max_episode_steps;
    v->o.continuous = continuous; v->o.enable_wind = enable_wind;
    v->o.gravity = gravity; v->o.wind_power = wind_power; v->o.turbulence_power = turbulence_power;
    v->w = (world_t *)calloc((size_t)n, sizeof(world_t));
    return v;
}
orc_lunar *orc_lunar_create(int64_t n, int max_episode_steps)
{
    return orc_lunar_create_ex(n, max_episode_steps, 0, 0, -10.0, 15.0, 1.5);
}
/* the two np.random.randint(-9999, 9999) draws of LunarLander.__init__ (:234-235), one pair per env */
void orc_lunar_set_wind_idx(orc_lunar *v, const int32_t *wind_idx, const int32_t *torque_idx)
{
    for (int64_t i = 0; i < v->n; i++) { v->w[i].wind_idx = wind_idx[i]; v->w[i].torque_idx = torque_idx[i]; }
}
void orc_lunar_get_wind_idx(const orc_lunar *v, int32_t *wind_idx, int32_t *torque_idx)
{
    for (int64_t i = 0; i < v->n; i++) { wind_idx[i] = v->w[i].wind_idx; torque_idx[i] = v->w[i].torque_idx; }
}
void orc_lunar_destroy(orc_lunar *v) { if (v) { free(v->w); free(v); } }

void orc_lunar_seed_range(orc_lunar *v, const uint32_t base[4], int64_t first)
{
    u128 b = 0;
    for (int k = 3; k >= 0; k--) b = (b << 32) | base[k];
    for (int64_t i = 0; i < v->n; i++) {
        u128 s = b + (u128)(uint64_t)(first + i);
        uint32_t ent[4] = {(uint32_t)s, (uint32_t)(s >> 32), (uint32_t)(s >> 64), (uint32_t)(s >> 96)};
        pcg_seed_from_words(&v->w[i].rng, ent);
    }
}

void orc_lunar_reset(orc_lunar *v, float *obs)
{
    for (int64_t i = 0; i < v->n; i++) lunar_reset_one(&v->w[i], &v->o, obs + 8 * i);
}

static void lunar_vec_one(orc_lunar *v, int64_t i, int action, const float *ca, float *obs, double *reward,
                          uint8_t *terminated, uint8_t *truncated, float *final_obs)
{
    world_t *W = &v->w[i];
    float o[8];
    double r;
    int term;
    lunar_step_one(W, &v->o, action, ca, o, &r, &term);
    W->elapsed += 1;
    int trunc = v->max_steps > 0 && W->elapsed >= v->max_steps;
    reward[i] = r;
    terminated[i] = (uint8_t)term;
    truncated[i] = (uint8_t)trunc;
    if (term || trunc) {
        if (final_obs) memcpy(final_obs + 8 * i, o, sizeof o);
        lunar_reset_one(W, &v->o, o);
    }
    memcpy(obs + 8 * i, o, sizeof o);
}

/* the envs are independent: a range of them per host thread (bench.py's cpu_baseline / reference arm) */
typedef struct {
    orc_lunar *v; const int64_t *ai; const float *af; float *obs; double *reward; uint8_t *terminated, *truncated;
    float *final_obs; int64_t lo, hi, invalid;
} lunar_job;

static void *lunar_range(void *arg)
{
    lunar_job *j = (lunar_job *)arg;
    const float zero[2] = {0.0f, 0.0f};
    for (int64_t i = j->lo; i < j->hi; i++) {
        if (j->af) { lunar_vec_one(j->v, i, 0, j->af + 2 * i, j->obs, j->reward, j->terminated, j->truncated, j->final_obs); continue; }
        if (j->ai[i] < 0 || j->ai[i] > 3) { j->invalid++; continue; }                       /* :482-484 */
        lunar_vec_one(j->v, i, (int)j->ai[i], zero, j->obs, j->reward, j->terminated, j->truncated, j->final_obs);
    }
    return NULL;
}

static int64_t lunar_run(orc_lunar *v, const int64_t *ai, const float *af, float *obs, double *reward, uint8_t *terminated,
                         uint8_t *truncated, float *final_obs, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > v->n) nthreads = (int)v->n;
    lunar_job jobs[256];
    pthread_t tid[256];
    const int64_t chunk = (v->n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        lunar_job *j = &jobs[t];
        j->v = v; j->ai = ai; j->af = af; j->obs = obs; j->reward = reward; j->terminated = terminated;
        j->truncated = truncated; j->final_obs = final_obs; j->invalid = 0;
        j->lo = t * chunk;
        j->hi = (j->lo + chunk < v->n) ? j->lo + chunk : v->n;
        if (j->lo > j->hi) j->lo = j->hi;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&tid[t], NULL, lunar_range, &jobs[t]);
    lunar_range(&jobs[0]);
    int64_t invalid = jobs[0].invalid;
    for (int t = 1; t < nthreads; t++) { pthread_join(tid[t], NULL); invalid += jobs[t].invalid; }
    return invalid;
}

int64_t orc_lunar_step_mt(orc_lunar *v, const int64_t *actions, float *obs, double *reward, uint8_t *terminated,
                          uint8_t *truncated, float *final_obs, int nthreads)
{
    return lunar_run(v, actions, NULL, obs, reward, terminated, truncated, final_obs, nthreads);
}

int64_t orc_lunar_step(orc_lunar *v, const int64_t *actions, float *obs, double *reward, uint8_t *terminated,
                       uint8_t *truncated, float *final_obs)
{
    return lunar_run(v, actions, NULL, obs, reward, terminated, truncated, final_obs, 1);
}

/* continuous=True: actions [n][2] float32, clipped to [-1, 1] inside the step (:480) */
void orc_lunar_step_cont_mt(orc_lunar *v, const float *actions, float *obs, double *reward, uint8_t *terminated,
                            uint8_t *truncated, float *final_obs, int nthreads)
{
    (void)lunar_run(v, NULL, actions, obs, reward, terminated, truncated, final_obs, nthreads);
}

void orc_lunar_step_cont(orc_lunar *v, const float *actions, float *obs, double *reward, uint8_t *terminated,
                         uint8_t *truncated, float *final_obs)
{
    (void)lunar_run(v, NULL, actions, obs, reward, terminated, truncated, final_obs, 1);
}

/* the 10 terrain edges of env i as 11 vertex heights (smooth_y, float32 as Box2D stores them) */
void orc_lunar_get_terrain(const orc_lunar *v, int64_t i, float *y11)
{
    const world_t *W = &v->w[i];
    for (int k = 0; k < 10; k++) y11[k] = W->e[k + 1].v1.y;
    y11[10] = W->e[10].v2.y;
}

/* workload statistics of the last step of every env: {touching contacts, position iterations} */
/* envs in which a touching pair was ever dropped because the manifold table was full */
int64_t orc_lunar_overflows(const orc_lunar *v)
{
    int64_t c = 0;
    for (int64_t i = 0; i < v->n; i++) c += v->w[i].overflows != 0;
    return c;
}

/* {b2TimeOfImpact evaluations, TOI sub-steps} summed over all envs since creation, {most sub-steps in one world step} */
void orc_lunar_toi_stats(const orc_lunar *v, int64_t out[3])
{
    out[0] = 0; out[1] = 0; out[2] = 0;
    for (int64_t i = 0; i < v->n; i++) {
        out[0] += v->w[i].toi_calls; out[1] += v->w[i].toi_events;
        if (v->w[i].toi_events_max > out[2]) out[2] = v->w[i].toi_events_max;
    }
}

void orc_lunar_get_stats(const orc_lunar *v, int32_t *out)
{
    for (int64_t i = 0; i < v->n; i++) { out[2 * i] = v->w[i].stat_contacts; out[2 * i + 1] = v->w[i].stat_pos_iters; }
}

/* test hook: b2TimeOfImpact (b2lite_toi.h) of a convex polygon (local vertices, body origin = centre of mass) sweeping
 * from (c0, a0) to (c1, a1) against the static edge v1-v2; returns the b2TOIOutput state (3 = e_touching,
 * 4 = e_separated, 2 = e_overlapped, 1 = e_failed) and writes t */
int orc_b2l_toi_probe(const float *poly_xy, int n, const float c0[2], float a0, const float c1[2], float a1,
                      const float v1[2], const float v2_[2], float *t_out)
{
    dproxy_t pA, pB;
    pA.count = 2; pA.v[0] = V(v1[0], v1[1]); pA.v[1] = V(v2_[0], v2_[1]);
    pB.count = n;
    for (int i = 0; i < n && i < MAXV; i++) pB.v[i] = V(poly_xy[2 * i], poly_xy[2 * i + 1]);
    sweep_t sA, sB;
    sA.localCenter = V(0.0f, 0.0f); sA.c0 = V(0.0f, 0.0f); sA.c = V(0.0f, 0.0f); sA.a0 = 0.0f; sA.a = 0.0f; sA.alpha0 = 0.0f;
    sB.localCenter = V(0.0f, 0.0f); sB.c0 = V(c0[0], c0[1]); sB.c = V(c1[0], c1[1]); sB.a0 = a0; sB.a = a1; sB.alpha0 = 0.0f;
    return time_of_impact(t_out, &pA, &sA, &pB, &sB, 1.0f);
}

/* test hook: b2Distance (GJK, b2lite_toi.h) between a convex polygon placed at (c, a) and the static edge v1-v2, core
 * shapes without radii, from an empty simplex cache; also returns the GJK iteration count through *cache_count */
float orc_b2l_distance_probe(const float *poly_xy, int n, const float c[2], float a, const float v1[2], const float v2_[2],
                             int *cache_count)
{
    dproxy_t pA, pB;
    pA.count = 2; pA.v[0] = V(v1[0], v1[1]); pA.v[1] = V(v2_[0], v2_[1]);
    pB.count = n;
    for (int i = 0; i < n && i < MAXV; i++) pB.v[i] = V(poly_xy[2 * i], poly_xy[2 * i + 1]);
    xform xfA = XF_ID, xfB;
    xfB.q = rot_of(a);
    xfB.p = V(c[0], c[1]);
    scache_t cache;
    cache.count = 0;
    float d = gjk_distance(&cache, &pA, xfA, &pB, xfB);
    if (cache_count) *cache_count = cache.count;
    return d;
}

/* test hook: overwrite the velocity of one body of env i (tunnelling tests) */
void orc_lunar_set_body_velocity(orc_lunar *v, int64_t i, int body, float vx, float vy, float w)
{
    world_t *W = &v->w[i];
    W->b[body].v = V(vx, vy);
    W->b[body].w = w;
}

/* debugging / parity: dump the 3 bodies (c.x, c.y, a, v.x, v.y, w) + flags of env i */
void orc_lunar_get_bodies(const orc_lunar *v, int64_t i, float out[18], int32_t flags[6])
{
    const world_t *W = &v->w[i];
    for (int b = 0; b < NB; b++) {
        out[6 * b + 0] = W->b[b].c.x; out[6 * b + 1] = W->b[b].c.y; out[6 * b + 2] = W->b[b].a;
        out[6 * b + 3] = W->b[b].v.x; out[6 * b + 4] = W->b[b].v.y; out[6 * b + 5] = W->b[b].w;
    }
    flags[0] = W->game_over; flags[1] = W->leg_contact[0]; flags[2] = W->leg_contact[1];
    flags[3] = W->b[0].awake; flags[4] = W->elapsed; flags[5] = 0;
    for (int k = 0; k < NB * NE; k++) flags[5] += W->ct[k].touching;
}
