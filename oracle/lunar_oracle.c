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
} world_t;

struct orc_lunar {
    int64_t n;
    int max_steps;
    world_t *w;
};

/* ContactDetector (lunar_lander.py:54-72) */
static void lunar_event(void *ctx, int body, int begin)
{
    world_t *W = (world_t *)ctx;
    if (begin) { if (body == 0) W->game_over = 1; else W->leg_contact[body - 1] = 1; }
    else if (body > 0) W->leg_contact[body - 1] = 0;
}

/* world.Step(1/50, 180, 60): island order {leg+1, lander, leg-1}, joint of leg+1 first */
static int world_step(world_t *W, float dt, int velIters, int posIters)
{
    static const int order[NB] = {2, 0, 1}, jorder[2] = {1, 0};
    b2l_world S;
    S.nb = NB; S.nj = 2; S.ne = NE;
    S.b = W->b; S.j = W->j; S.e = W->e; S.ct = W->ct;
    S.body_order = order; S.joint_order = jorder;
    S.inv_dt0 = W->inv_dt0;
    S.event = lunar_event; S.ctx = W;
    b2l_step(&S, dt, velIters, posIters);
    W->inv_dt0 = S.inv_dt0;
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

static void lunar_step_one(world_t *W, int action, float *obs, double *reward, int *terminated);

/* lunar_lander.py:308-420 */
static void lunar_reset_one(world_t *W, float *obs)
{
    pcg64_t rng = W->rng;
    int32_t elapsed_dummy = 0;
    float inv_dt0 = W->inv_dt0; /* the b2World object survives reset() */
    memset(W, 0, sizeof *W);
    W->inv_dt0 = inv_dt0;
    (void)elapsed_dummy;
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
    lunar_step_one(W, 0, obs, &r, &t);                                                     /* :420 */
}

/* lunar_lander.py:444-600 (discrete actions, no wind) */
static void lunar_step_one(world_t *W, int action, float *obs, double *reward, int *terminated)
{
    body_t *L = &W->b[0];
    double ang = (double)L->a;
    double tip0 = sin(ang), tip1 = cos(ang);                                               /* :487 */
    double side0 = -tip1, side1 = tip0;
    double disp0 = rng_uniform(&W->rng, -1.0, +1.0) / SCALE;                               /* :489 */
    double disp1 = rng_uniform(&W->rng, -1.0, +1.0) / SCALE;
    double m_power = 0.0;
    if (action == 2) {                                                                     /* :491-520 */
        m_power = 1.0;
        double ox = tip0 * (4 / SCALE + 2 * disp0) + side0 * disp1;
        double oy = -tip1 * (4 / SCALE + 2 * disp0) - side1 * disp1;
        double px = (double)L->xf.p.x + ox, py = (double)L->xf.p.y + oy;
        v2 imp = V((float)(-ox * MAIN_ENGINE_POWER * m_power), (float)(-oy * MAIN_ENGINE_POWER * m_power));
        v2 pt = V((float)px, (float)py);
        if (L->awake) { /* b2Body::ApplyLinearImpulse */
            L->v = add(L->v, scl(L->invMass, imp));
            L->w += L->invI * crs(sub(pt, L->c), imp);
        }
    }
    double s_power = 0.0;
    if (action == 1 || action == 3) {                                                      /* :522-554 */
        double direction = action - 2;
        s_power = 1.0;
        double ox = tip0 * disp0 + side0 * (3 * disp1 + direction * SIDE_ENGINE_AWAY / SCALE);
        double oy = -tip1 * disp0 - side1 * (3 * disp1 + direction * SIDE_ENGINE_AWAY / SCALE);
        double px = (double)L->xf.p.x + ox - tip0 * 17 / SCALE;
        double py = (double)L->xf.p.y + oy + tip1 * SIDE_ENGINE_HEIGHT / SCALE;
        v2 imp = V((float)(-ox * SIDE_ENGINE_POWER * s_power), (float)(-oy * SIDE_ENGINE_POWER * s_power));
        v2 pt = V((float)px, (float)py);
        if (L->awake) {
            L->v = add(L->v, scl(L->invMass, imp));
            L->w += L->invI * crs(sub(pt, L->c), imp);
        }
    }
    int awake = world_step(W, (float)(1.0 / FPS), 6 * 30, 2 * 30);                         /* :556 */
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
    r -= m_power * 0.30;                                                                   /* :585-588 */
    r -= s_power * 0.03;
    int term = 0;
    if (W->game_over || fabs(st[0]) >= 1.0) { term = 1; r = -100; }                        /* :590-593 */
    if (!awake) { term = 1; r = +100; }                                                 /* :594-596 */
    for (int k = 0; k < 8; k++) obs[k] = (float)st[k];                                     /* :600 */
    *reward = r;
    *terminated = term;
}

/* ---------------------------------------------------------------- vector API */
orc_lunar *orc_lunar_create(int64_t n, int max_episode_steps)
{
    if (n <= 0) return NULL;
    orc_lunar *v = (orc_lunar *)calloc(1, sizeof *v);
    v->n = n;
    v->max_steps = max_episode_steps;
    v->w = (world_t *)calloc((size_t)n, sizeof(world_t));
    return v;
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
    for (int64_t i = 0; i < v->n; i++) lunar_reset_one(&v->w[i], obs + 8 * i);
}

int64_t orc_lunar_step(orc_lunar *v, const int64_t *actions, float *obs, double *reward, uint8_t *terminated,
                       uint8_t *truncated, float *final_obs)
{
    int64_t invalid = 0;
    for (int64_t i = 0; i < v->n; i++) {
        world_t *W = &v->w[i];
        if (actions[i] < 0 || actions[i] > 3) { invalid++; continue; }                     /* :482-484 */
        float o[8];
        double r;
        int term;
        lunar_step_one(W, (int)actions[i], o, &r, &term);
        W->elapsed += 1;
        int trunc = v->max_steps > 0 && W->elapsed >= v->max_steps;
        reward[i] = r;
        terminated[i] = (uint8_t)term;
        truncated[i] = (uint8_t)trunc;
        if (term || trunc) {
            if (final_obs) memcpy(final_obs + 8 * i, o, sizeof o);
            lunar_reset_one(W, o);
        }
        memcpy(obs + 8 * i, o, sizeof o);
    }
    return invalid;
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
