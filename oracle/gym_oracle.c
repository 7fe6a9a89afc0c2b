/*
 * gym_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of openai/gym 0.26.2's classic-control step()/reset()
 * path, written so that every floating-point operation happens in the same
 * precision and order as the reference evaluated under numpy >= 2 (NEP 50
 * promotion).  Build with -ffp-contract=off (see oracle/Makefile): the
 * reference rounds after every multiply and add.
 *
 * Where the reference writes `x ** 2` on a Python float / numpy scalar it ends
 * up in libm pow()/powf(); this file calls the same libm function so that the
 * two agree to the last bit on the same glibc.
 *
 * Parity status: PINNED against the reference (oracle/gen_golden.py) and the
 * fixtures in tests/golden/.
 */
#include "gym_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------ */
/* numpy bit generator: PCG64 (XSL-RR 128/64) + SeedSequence                 */
/* reference call sites: gym/utils/seeding.py:24-26; algorithm: numpy        */
/* (third-party, not vendored; restated from its published description,      */
/* SURVEY.md Appendix C; checked against numpy 2.3.5 in tests).              */
/* ------------------------------------------------------------------------ */

typedef struct {
    u128 state;
    u128 inc;
} pcg64_t;

#define PCG_MULT ((((u128)0x2360ED051FC65DA4ULL) << 64) | (u128)0x4385DF649FCCF645ULL)

static inline void pcg64_advance(pcg64_t *g) { g->state = g->state * PCG_MULT + g->inc; }

static void pcg64_seed(pcg64_t *g, u128 initstate, u128 initseq)
{
    g->state = 0;
    g->inc = (initseq << 1) | 1;
    pcg64_advance(g);
    g->state += initstate;
    pcg64_advance(g);
}

static inline uint64_t pcg64_next64(pcg64_t *g)
{
    pcg64_advance(g);
    uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((-rot) & 63));
}

static inline double pcg64_next_double(pcg64_t *g)
{
    return (double)(pcg64_next64(g) >> 11) * (1.0 / 9007199254740992.0);
}

/* Generator.uniform(low, high): low + (high - low) * next_double */
static inline double rng_uniform(pcg64_t *g, double low, double high)
{
    double range = high - low;
    return low + range * pcg64_next_double(g);
}

void orc_seed_sequence(const uint32_t ent[4], uint64_t out[4])
{
    const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u;
    const uint32_t INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
    const uint32_t MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
    uint32_t hc = INIT_A, pool[4];
#define HASHMIX(dst, val)                 \
    do {                                  \
        uint32_t v_ = (val);              \
        v_ ^= hc;                         \
        hc *= MULT_A;                     \
        v_ *= hc;                         \
        v_ ^= v_ >> 16;                   \
        (dst) = v_;                       \
    } while (0)
    for (int i = 0; i < 4; i++) HASHMIX(pool[i], ent[i]);
    for (int src = 0; src < 4; src++)
        for (int dst = 0; dst < 4; dst++)
            if (src != dst) {
                uint32_t h;
                HASHMIX(h, pool[src]);
                uint32_t r = MIX_L * pool[dst] - MIX_R * h;
                r ^= r >> 16;
                pool[dst] = r;
            }
#undef HASHMIX
    uint32_t o[8];
    hc = INIT_B;
    for (int i = 0; i < 8; i++) {
        uint32_t v = pool[i & 3];
        v ^= hc;
        hc *= MULT_B;
        v *= hc;
        v ^= v >> 16;
        o[i] = v;
    }
    for (int k = 0; k < 4; k++) out[k] = (uint64_t)o[2 * k] | ((uint64_t)o[2 * k + 1] << 32);
}

/* ------------------------------------------------------------------------ */
/* vector container                                                          */
/* ------------------------------------------------------------------------ */

#define MAX_STATE 4
#define MAX_OBS 6

struct orc_vec {
    int kind;
    int64_t n;
    int max_steps;
    double param0;
    double *state;    /* [n][state_dim] */
    int32_t *elapsed; /* [n] TimeLimit._elapsed_steps */
    pcg64_t *rng;     /* [n] */
};

static const int k_obs_dim[ORC_NUM_KINDS] = {4, 2, 2, 3, 6};
static const int k_act_dim[ORC_NUM_KINDS] = {0, 0, 1, 1, 0};
static const int k_state_dim[ORC_NUM_KINDS] = {4, 2, 2, 2, 4};
static const int k_num_actions[ORC_NUM_KINDS] = {2, 3, 0, 0, 3};

int orc_obs_dim(int kind) { return k_obs_dim[kind]; }
int orc_act_dim(int kind) { return k_act_dim[kind]; }
int orc_state_dim(int kind) { return k_state_dim[kind]; }
int orc_num_actions(int kind) { return k_num_actions[kind]; }

orc_vec *orc_vec_create(int kind, int64_t n, int max_episode_steps, double param0)
{
    if (kind < 0 || kind >= ORC_NUM_KINDS || n <= 0) return NULL;
    orc_vec *v = (orc_vec *)calloc(1, sizeof(*v));
    v->kind = kind;
    v->n = n;
    v->max_steps = max_episode_steps;
    v->param0 = param0;
    v->state = (double *)calloc((size_t)n * k_state_dim[kind], sizeof(double));
    v->elapsed = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    v->rng = (pcg64_t *)calloc((size_t)n, sizeof(pcg64_t));
    return v;
}

void orc_vec_destroy(orc_vec *v)
{
    if (!v) return;
    free(v->state);
    free(v->elapsed);
    free(v->rng);
    free(v);
}

void orc_vec_seed_env(orc_vec *v, int64_t i, const uint32_t ent[4])
{
    uint64_t w[4];
    orc_seed_sequence(ent, w);
    pcg64_seed(&v->rng[i], ((u128)w[0] << 64) | w[1], ((u128)w[2] << 64) | w[3]);
}

void orc_vec_seed_range(orc_vec *v, const uint32_t base[4], int64_t first)
{
    u128 b = 0;
    for (int k = 3; k >= 0; k--) b = (b << 32) | base[k];
    for (int64_t i = 0; i < v->n; i++) {
        u128 s = b + (u128)(uint64_t)(first + i);
        uint32_t ent[4] = {(uint32_t)s, (uint32_t)(s >> 32), (uint32_t)(s >> 64), (uint32_t)(s >> 96)};
        orc_vec_seed_env(v, i, ent);
    }
}

void orc_vec_get_rng(const orc_vec *v, uint64_t *out)
{
    for (int64_t i = 0; i < v->n; i++) {
        out[4 * i + 0] = (uint64_t)(v->rng[i].state >> 64);
        out[4 * i + 1] = (uint64_t)v->rng[i].state;
        out[4 * i + 2] = (uint64_t)(v->rng[i].inc >> 64);
        out[4 * i + 3] = (uint64_t)v->rng[i].inc;
    }
}

void orc_vec_set_rng(orc_vec *v, const uint64_t *in)
{
    for (int64_t i = 0; i < v->n; i++) {
        v->rng[i].state = ((u128)in[4 * i + 0] << 64) | in[4 * i + 1];
        v->rng[i].inc = ((u128)in[4 * i + 2] << 64) | in[4 * i + 3];
    }
}

double orc_vec_next_double(orc_vec *v, int64_t i) { return pcg64_next_double(&v->rng[i]); }

void orc_vec_get_state(const orc_vec *v, double *state, int32_t *elapsed)
{
    if (state) memcpy(state, v->state, sizeof(double) * (size_t)v->n * k_state_dim[v->kind]);
    if (elapsed) memcpy(elapsed, v->elapsed, sizeof(int32_t) * (size_t)v->n);
}

void orc_vec_set_state(orc_vec *v, const double *state, const int32_t *elapsed)
{
    if (state) memcpy(v->state, state, sizeof(double) * (size_t)v->n * k_state_dim[v->kind]);
    if (elapsed) memcpy(v->elapsed, elapsed, sizeof(int32_t) * (size_t)v->n);
}

/* ------------------------------------------------------------------------ */
/* CartPole -- gym/envs/classic_control/cartpole.py                          */
/* ------------------------------------------------------------------------ */

/* cartpole.py:190-207 */
static void cartpole_reset(double *s, pcg64_t *g, double low, double high, float *obs)
{
    for (int k = 0; k < 4; k++) s[k] = rng_uniform(g, low, high); /* :202 */
    for (int k = 0; k < 4; k++) obs[k] = (float)s[k];             /* :207 */
}

/* cartpole.py:130-188 (euler integrator branch :149-153) */
static void cartpole_step(double *s, int64_t action, float *obs, double *reward, int *terminated)
{
    const double gravity = 9.8, masscart = 1.0, masspole = 0.1;   /* :90-92 */
    const double total_mass = masspole + masscart;                /* :93 */
    const double length = 0.5;                                    /* :94 */
    const double polemass_length = masspole * length;             /* :95 */
    const double force_mag = 10.0, tau = 0.02;                    /* :96-97 */
    const double theta_threshold = 12 * 2 * M_PI / 360;           /* :101 */
    const double x_threshold = 2.4;                               /* :102 */

    double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    double force = (action == 1) ? force_mag : -force_mag;        /* :135 */
    double costheta = cos(theta), sintheta = sin(theta);          /* :136-137 */
    double temp = (force + polemass_length * pow(theta_dot, 2.0) * sintheta) / total_mass; /* :141-143 */
    double thetaacc = (gravity * sintheta - costheta * temp) /
                      (length * (4.0 / 3.0 - masspole * pow(costheta, 2.0) / total_mass)); /* :144-146 */
    double xacc = temp - polemass_length * thetaacc * costheta / total_mass;               /* :147 */
    x = x + tau * x_dot;                                          /* :150 */
    x_dot = x_dot + tau * xacc;
    theta = theta + tau * theta_dot;
    theta_dot = theta_dot + tau * thetaacc;
    s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;       /* :160 */
    *terminated = (x < -x_threshold || x > x_threshold || theta < -theta_threshold ||
                   theta > theta_threshold);                      /* :162-167 */
    *reward = 1.0;                                                /* :169-174 (autoreset: never beyond) */
    for (int k = 0; k < 4; k++) obs[k] = (float)s[k];             /* :188 */
}

/* ------------------------------------------------------------------------ */
/* MountainCar-v0 -- gym/envs/classic_control/mountain_car.py                */
/* ------------------------------------------------------------------------ */

static inline double clip_f64(double x, double lo, double hi)
{ /* np.clip == minimum(maximum(x, lo), hi) */
    double m = (x < lo) ? lo : x;
    return (m > hi) ? hi : m;
}

/* mountain_car.py:150-164 (identical in continuous_mountain_car.py:177-186) */
static void mountaincar_reset(double *s, pcg64_t *g, double low, double high, float *obs)
{
    s[0] = rng_uniform(g, low, high); /* :160 */
    s[1] = 0.0;
    obs[0] = (float)s[0];
    obs[1] = (float)s[1];
}

/* mountain_car.py:127-148 */
static void mountaincar_step(double *s, int64_t action, double goal_velocity, float *obs,
                             double *reward, int *terminated)
{
    const double min_position = -1.2, max_position = 0.6, max_speed = 0.07;
    const double goal_position = 0.5, force = 0.001, gravity = 0.0025; /* :104-111 */
    double position = s[0], velocity = s[1];
    velocity += (double)(action - 1) * force + cos(3 * position) * (-gravity); /* :132 */
    velocity = clip_f64(velocity, -max_speed, max_speed);                      /* :133 */
    position += velocity;                                                      /* :134 */
    position = clip_f64(position, min_position, max_position);                 /* :135 */
    if (position == min_position && velocity < 0) velocity = 0;                /* :136-137 */
    *terminated = (position >= goal_position && velocity >= goal_velocity);    /* :139-141 */
    *reward = -1.0;                                                            /* :142 */
    s[0] = position; s[1] = velocity;                                          /* :144 */
    obs[0] = (float)position; obs[1] = (float)velocity;                        /* :148 */
}

/* ------------------------------------------------------------------------ */
/* MountainCarContinuous-v0 -- continuous_mountain_car.py:142-175            */
/* Under numpy >= 2 the state is an f64 array only between reset (:182) and  */
/* the first step; afterwards it is the f32 array built at :171, so          */
/* the arithmetic runs in float32 with Python-float "weak" constants.        */
/* `fresh` says which of the two dtypes position/velocity currently have.    */
/* ------------------------------------------------------------------------ */
static void mcc_step(double *s, int fresh, float a0, double goal_velocity, float *obs,
                     double *reward, int *terminated)
{
    const double min_action = -1.0, max_action = 1.0, min_position = -1.2, max_position = 0.6;
    const double max_speed = 0.07, goal_position = 0.45, power = 0.0015; /* :110-119 */
    /* T(x): round a Python float to the dtype of the numpy scalar it meets */
#define T(x) (fresh ? (double)(x) : (double)(float)(x))
    double position = s[0], velocity = s[1]; /* :144-145 (dtype f64 if fresh else f32) */
    int vel_py = 0, pos_py = 0;              /* value became a plain Python number */

    /* :146 force = min(max(action[0], -1.0), 1.0): the Python constant is
     * returned only when it strictly clips */
    int force_py = 0;
    double force_c = 0.0;
    if (a0 < (float)min_action) { force_py = 1; force_c = min_action; }
    else if (a0 > (float)max_action) { force_py = 1; force_c = max_action; }

    /* :148 velocity += force * power - 0.0025 * math.cos(3 * position) */
    double three_p = fresh ? 3 * position : (double)((float)3 * (float)position);
    double c = 0.0025 * cos(three_p);
    if (force_py) {
        double inc = force_c * power - c;              /* Python floats */
        velocity = fresh ? velocity + inc : (double)((float)velocity + (float)inc);
    } else {
        float inc = a0 * (float)power - (float)c;      /* np.float32 arithmetic */
        velocity = fresh ? velocity + (double)inc : (double)((float)velocity + inc);
    }
    if (velocity > T(max_speed)) { velocity = max_speed; vel_py = 1; }                 /* :149-150 */
    if (vel_py ? (velocity < -max_speed) : (velocity < T(-max_speed))) {               /* :151-152 */
        velocity = -max_speed; vel_py = 1;
    }
    /* :153 position += velocity (result keeps position's dtype) */
    position = fresh ? position + velocity
                     : (double)((float)position + (float)velocity);
    if (position > T(max_position)) { position = max_position; pos_py = 1; }           /* :154-155 */
    if (pos_py ? (position < min_position) : (position < T(min_position))) {           /* :156-157 */
        position = min_position; pos_py = 1;
    }
    int at_min = pos_py ? (position == min_position) : (position == T(min_position));
    int vneg = vel_py ? (velocity < 0) : (velocity < 0);
    if (at_min && vneg) { velocity = 0; vel_py = 1; }                                  /* :158-159 */

    int pos_ok = pos_py ? (position >= goal_position) : (position >= T(goal_position));
    int vel_ok = vel_py ? (velocity >= goal_velocity) : (velocity >= T(goal_velocity));
    *terminated = pos_ok && vel_ok;                                                    /* :162-164 */
    double r = 0;
    if (*terminated) r = 100.0;                                                        /* :166-168 */
    r -= pow((double)a0, 2.0) * 0.1;                                                   /* :169 */
    *reward = r;
    s[0] = (double)(float)position; /* :171 np.array([...], dtype=np.float32) */
    s[1] = (double)(float)velocity;
    obs[0] = (float)position; obs[1] = (float)velocity;                                /* :175 */
#undef T
}

/* ------------------------------------------------------------------------ */
/* Pendulum-v1 -- gym/envs/classic_control/pendulum.py                       */
/* ------------------------------------------------------------------------ */

/* Python-style float modulo as numpy implements `%` on float64 scalars */
static inline double py_mod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* pendulum.py:270-271 */
static inline double angle_normalize(double x) { return py_mod(x + M_PI, 2 * M_PI) - M_PI; }

/* pendulum.py:161-163 */
static void pendulum_obs(const double *s, float *obs)
{
    obs[0] = (float)cos(s[0]);
    obs[1] = (float)sin(s[0]);
    obs[2] = (float)s[1];
}

/* pendulum.py:141-159 */
static void pendulum_reset(double *s, pcg64_t *g, double x_init, double y_init, float *obs)
{
    s[0] = rng_uniform(g, -x_init, x_init); /* :153-154 vector low/high, C order */
    s[1] = rng_uniform(g, -y_init, y_init);
    pendulum_obs(s, obs);
}

/* pendulum.py:119-139 */
static void pendulum_step(double *s, float a0, double g, float *obs, double *reward)
{
    const double max_speed = 8, max_torque = 2.0, dt = 0.05, m = 1.0, l = 1.0; /* :96-101 */
    double th = s[0], thdot = s[1];
    /* :127 u = np.clip(u, -2, 2)[0]  (np.float32) */
    float u = a0;
    u = (u < (float)-max_torque) ? (float)-max_torque : u;
    u = (u > (float)max_torque) ? (float)max_torque : u;
    /* :129 costs: the u-term is float32 (u**2 -> powf, 0.001 weak) */
    float u_term = (float)0.001 * powf(u, 2.0f);
    double costs = pow(angle_normalize(th), 2.0) + 0.1 * pow(thdot, 2.0) + (double)u_term;
    /* :131 3.0/(m*l**2)*u is float32, promoted on the add */
    float tq = (float)(3.0 / (m * pow(l, 2.0))) * u;
    double newthdot = thdot + (3 * g / (2 * l) * sin(th) + (double)tq) * dt;
    newthdot = clip_f64(newthdot, -max_speed, max_speed); /* :132 */
    double newth = th + newthdot * dt;                    /* :133 */
    s[0] = newth; s[1] = newthdot;                        /* :135 */
    pendulum_obs(s, obs);
    *reward = -costs;                                     /* :139 */
}

/* ------------------------------------------------------------------------ */
/* Acrobot-v1 -- gym/envs/classic_control/acrobot.py                         */
/* ------------------------------------------------------------------------ */

/* acrobot.py:237-277 ("book" branch :271-276) */
static void acrobot_dsdt(const double *sa, double *out)
{
    const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 0.5, I1 = 1.0, I2 = 1.0; /* :145-151 */
    const double g = 9.8;                                                               /* :245 */
    double a = sa[4];
    double theta1 = sa[0], theta2 = sa[1], dtheta1 = sa[2], dtheta2 = sa[3];
    double d1 = m1 * pow(lc1, 2.0) + m2 * (pow(l1, 2.0) + pow(lc2, 2.0) + 2 * l1 * lc2 * cos(theta2)) + I1 + I2;
    double d2 = m2 * (pow(lc2, 2.0) + l1 * lc2 * cos(theta2)) + I2;
    double phi2 = m2 * lc2 * g * cos(theta1 + theta2 - M_PI / 2.0);
    double phi1 = -m2 * l1 * lc2 * pow(dtheta2, 2.0) * sin(theta2)
                  - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2)
                  + (m1 * lc1 + m2 * l1) * g * cos(theta1 - M_PI / 2)
                  + phi2;
    double ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * pow(dtheta1, 2.0) * sin(theta2) - phi2) /
                      (m2 * pow(lc2, 2.0) + I2 - pow(d2, 2.0) / d1);
    double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
    out[0] = dtheta1; out[1] = dtheta2; out[2] = ddtheta1; out[3] = ddtheta2; out[4] = 0.0;
}

/* acrobot.py:378-396 */
static double acrobot_wrap(double x, double m, double M)
{
    double diff = M - m;
    while (x > M) x = x - diff;
    while (x < m) x = x + diff;
    return x;
}

/* acrobot.py:399-415 bound(x, m, M) = min(max(x, m), M) */
static double acrobot_bound(double x, double m, double M)
{
    double t = (m > x) ? m : x;   /* max(x, m): returns m only if m > x */
    return (M < t) ? M : t;       /* min(t, M): returns M only if M < t */
}

/* acrobot.py:225-230 with an f64 state (after a step) */
static void acrobot_obs64(const double *s, float *obs)
{
    obs[0] = (float)cos(s[0]); obs[1] = (float)sin(s[0]);
    obs[2] = (float)cos(s[1]); obs[3] = (float)sin(s[1]);
    obs[4] = (float)s[2]; obs[5] = (float)s[3];
}

/* acrobot.py:181-194: the state is cast to float32 (:188-190), so _get_ob
 * evaluates numpy's own float32 cos/sin kernels on it.  Those are neither
 * libm's cosf/sinf nor correctly rounded (numpy documents <= ~1.5 ulp); their
 * source is not in /root/reference.  The oracle returns the correctly rounded
 * float32 value (f64 libm, then one rounding): measured against numpy 2.3.5 on
 * U(-0.1, 0.1) it is bit-identical for 99.9 % of inputs and 1 float32 ulp off
 * otherwise.  These are reset observations only -- they never feed the
 * dynamics, which restart from the float32-rounded state (:188-190). */
static void acrobot_reset(double *s, pcg64_t *g, double low, double high, float *obs)
{
    for (int k = 0; k < 4; k++) s[k] = (double)(float)rng_uniform(g, low, high);
    obs[0] = (float)cos(s[0]); obs[1] = (float)sin(s[0]);
    obs[2] = (float)cos(s[1]); obs[3] = (float)sin(s[1]);
    obs[4] = (float)s[2]; obs[5] = (float)s[3];
}

/* acrobot.py:196-223 + rk4 :418-465 over t = [0, 0.2] */
static void acrobot_step(double *s, int64_t action, float *obs, double *reward, int *terminated)
{
    static const double AVAIL_TORQUE[3] = {-1.0, 0.0, +1};  /* :156 */
    const double MAX_VEL_1 = 4 * M_PI, MAX_VEL_2 = 9 * M_PI; /* :153-154 */
    const double dt = 0.2 - 0, dt2 = dt / 2.0;               /* :453-455 */
    double y0[5] = {s[0], s[1], s[2], s[3], AVAIL_TORQUE[action]}; /* :207 */
    double k1[5], k2[5], k3[5], k4[5], y[5];
    acrobot_dsdt(y0, k1);                                    /* :458 */
    for (int i = 0; i < 5; i++) y[i] = y0[i] + dt2 * k1[i];
    acrobot_dsdt(y, k2);                                     /* :459 */
    for (int i = 0; i < 5; i++) y[i] = y0[i] + dt2 * k2[i];
    acrobot_dsdt(y, k3);                                     /* :460 */
    for (int i = 0; i < 5; i++) y[i] = y0[i] + dt * k3[i];
    acrobot_dsdt(y, k4);                                     /* :461 */
    double ns[4];
    for (int i = 0; i < 4; i++)                              /* :462 */
        ns[i] = y0[i] + dt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    ns[0] = acrobot_wrap(ns[0], -M_PI, M_PI);                /* :213-214 */
    ns[1] = acrobot_wrap(ns[1], -M_PI, M_PI);
    ns[2] = acrobot_bound(ns[2], -MAX_VEL_1, MAX_VEL_1);     /* :215-216 */
    ns[3] = acrobot_bound(ns[3], -MAX_VEL_2, MAX_VEL_2);
    for (int i = 0; i < 4; i++) s[i] = ns[i];                /* :217 */
    *terminated = (-cos(s[0]) - cos(s[1] + s[0]) > 1.0);     /* :235 */
    *reward = *terminated ? 0.0 : -1.0;                      /* :219 */
    acrobot_obs64(s, obs);
}

/* ------------------------------------------------------------------------ */
/* vector reset / step                                                       */
/* ------------------------------------------------------------------------ */

static void default_bounds(int kind, double *b)
{
    switch (kind) {
    case ORC_CARTPOLE: b[0] = -0.05; b[1] = 0.05; break;              /* cartpole.py:199-201 */
    case ORC_MOUNTAINCAR:
    case ORC_MOUNTAINCAR_CONT: b[0] = -0.6; b[1] = -0.4; break;       /* mountain_car.py:159 */
    case ORC_PENDULUM: b[0] = M_PI; b[1] = 1.0; break;                /* pendulum.py:14-15 */
    case ORC_ACROBOT: b[0] = -0.1; b[1] = 0.1; break;                 /* acrobot.py:185-187 */
    }
}

static void reset_one(orc_vec *v, int64_t i, const double *b, float *obs)
{
    double *s = v->state + i * k_state_dim[v->kind];
    pcg64_t *g = &v->rng[i];
    switch (v->kind) {
    case ORC_CARTPOLE: cartpole_reset(s, g, b[0], b[1], obs); break;
    case ORC_MOUNTAINCAR:
    case ORC_MOUNTAINCAR_CONT: mountaincar_reset(s, g, b[0], b[1], obs); break;
    case ORC_PENDULUM: pendulum_reset(s, g, b[0], b[1], obs); break;
    case ORC_ACROBOT: acrobot_reset(s, g, b[0], b[1], obs); break;
    }
    v->elapsed[i] = 0; /* time_limit.py:67 */
}

void orc_vec_reset(orc_vec *v, const uint8_t *mask, const double *bounds, float *obs)
{
    double b[2];
    if (bounds) { b[0] = bounds[0]; b[1] = bounds[1]; }
    else default_bounds(v->kind, b);
    const int D = k_obs_dim[v->kind];
    for (int64_t i = 0; i < v->n; i++)
        if (!mask || mask[i]) reset_one(v, i, b, obs + i * D);
}

typedef struct {
    orc_vec *v;
    const void *actions;
    float *obs;
    double *reward;
    uint8_t *terminated, *truncated;
    float *final_obs;
    int64_t lo, hi, invalid;
} step_job;

static void *step_range(void *arg)
{
    step_job *j = (step_job *)arg;
    orc_vec *v = j->v;
    const int kind = v->kind, D = k_obs_dim[kind], S = k_state_dim[kind];
    const int64_t *ai = (const int64_t *)j->actions;
    const float *af = (const float *)j->actions;
    double defb[2];
    default_bounds(kind, defb);
    j->invalid = 0;
    for (int64_t i = j->lo; i < j->hi; i++) {
        double *s = v->state + i * S;
        float o[MAX_OBS];
        double r = 0.0;
        int term = 0;
        if (k_act_dim[kind] == 0 && (ai[i] < 0 || ai[i] >= k_num_actions[kind])) {
            j->invalid++; /* cartpole.py:132, mountain_car.py:128-130, acrobot.py:199 */
            continue;
        }
        switch (kind) {
        case ORC_CARTPOLE: cartpole_step(s, ai[i], o, &r, &term); break;
        case ORC_MOUNTAINCAR: mountaincar_step(s, ai[i], v->param0, o, &r, &term); break;
        case ORC_MOUNTAINCAR_CONT: mcc_step(s, v->elapsed[i] == 0, af[i], v->param0, o, &r, &term); break;
        case ORC_PENDULUM: pendulum_step(s, af[i], v->param0, o, &r); break;
        case ORC_ACROBOT: acrobot_step(s, ai[i], o, &r, &term); break;
        }
        v->elapsed[i] += 1;                                                /* time_limit.py:51 */
        int trunc = (v->max_steps > 0 && v->elapsed[i] >= v->max_steps);   /* :53-54 */
        j->reward[i] = r;
        j->terminated[i] = (uint8_t)term;
        j->truncated[i] = (uint8_t)trunc;
        if (term || trunc) {                                               /* sync_vector_env.py:152-156 */
            if (j->final_obs) memcpy(j->final_obs + i * D, o, sizeof(float) * D);
            reset_one(v, i, defb, o);
        }
        memcpy(j->obs + i * D, o, sizeof(float) * D);
    }
    return NULL;
}

int64_t orc_vec_step(orc_vec *v, const void *actions, float *obs, double *reward,
                     uint8_t *terminated, uint8_t *truncated, float *final_obs, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > v->n) nthreads = (int)v->n;
    step_job jobs[256];
    pthread_t tid[256];
    int64_t chunk = (v->n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        step_job *j = &jobs[t];
        j->v = v; j->actions = actions; j->obs = obs; j->reward = reward;
        j->terminated = terminated; j->truncated = truncated; j->final_obs = final_obs;
        j->lo = t * chunk;
        j->hi = (j->lo + chunk < v->n) ? j->lo + chunk : v->n;
        if (j->lo > j->hi) j->lo = j->hi;
        j->invalid = 0;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&tid[t], NULL, step_range, &jobs[t]);
    step_range(&jobs[0]);
    int64_t invalid = jobs[0].invalid;
    for (int t = 1; t < nthreads; t++) {
        pthread_join(tid[t], NULL);
        invalid += jobs[t].invalid;
    }
    return invalid;
}
