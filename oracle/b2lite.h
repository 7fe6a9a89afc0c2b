/*
 * b2lite.h -- CPU ORACLE building block (test infrastructure, NOT product code).
 *
 * The slice of Box2D 2.3's algorithm that the Box2D gym tasks exercise, restated from its published
 * design (Erin Catto's sequential-impulse solver as shipped in Box2D 2.3; box2d-py == 2.3.5,
 * setup.py:15, is neither vendored in /root/reference nor installable here): convex polygons against
 * static edges, revolute joints with limit + motor, one island per scene.
 *   b2World::Step(dt, velIters, posIters): Collide (edge-vs-polygon manifolds, begin/end events,
 *   warm-start impulse matching by feature id) -> island solve (integrate velocities, joints, 2-point
 *   block contact solver with friction, position integration with translation clamps, Baumgarte
 *   position iterations with early exit) -> sleeping.
 *   -> SolveTOI (b2lite_toi.h: b2TimeOfImpact + TOI sub-steps of dynamic bodies against static fixtures).
 * PARITY UNPINNED.  Deliberate deviations from Box2D: the broad phase is replaced by testing every
 * (body, edge) pair behind the same fat AABB reject; constraints are ordered by the scene's island order x descending edge index; sin/cos
 * of body angles use the polynomial below instead of the platform libm.
 * All arithmetic is IEEE float32 with separate rounding of every operation (-ffp-contract=off).
 */
#ifndef B2LITE_H
#define B2LITE_H
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ---------------------------------------------------------------- float32 vector helpers */
typedef struct { float x, y; } v2;
typedef struct { float s, c; } rot;
typedef struct { v2 p; rot q; } xform;

static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 add(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 sub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 neg(v2 a) { return V(-a.x, -a.y); }
static inline v2 scl(float s, v2 a) { return V(s * a.x, s * a.y); }
static inline float dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float crs(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline v2 crs_vs(v2 a, float s) { return V(s * a.y, -s * a.x); }  /* b2Cross(vec, scalar) */
static inline v2 crs_sv(float s, v2 a) { return V(-s * a.y, s * a.x); }  /* b2Cross(scalar, vec) */
static inline v2 rmul(rot q, v2 v) { return V(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
static inline v2 rmulT(rot q, v2 v) { return V(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
static inline v2 xmul(xform T, v2 v) { return V((T.q.c * v.x - T.q.s * v.y) + T.p.x, (T.q.s * v.x + T.q.c * v.y) + T.p.y); }
static inline v2 xmulT(xform T, v2 v) { float px = v.x - T.p.x, py = v.y - T.p.y; return V(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py); }
static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float clampf(float a, float lo, float hi) { return fmaxf_(lo, fminf_(a, hi)); }

/* sin/cos of a body angle: Cody-Waite reduction by pi/2 (two fmaf) + cephes-style minimax
 * polynomials on [-pi/4, pi/4]; ~1 ulp.  The CUDA kernel evaluates the same sequence. */
static void lite_sincosf(float x, float *sn, float *cs)
{
    float kf = rintf(x * 0.636619772367581343f);
    int k = (int)kf;
    float r = fmaf(kf, -1.5707963705062866f, x);
    r = fmaf(kf, 4.371139000186241e-08f, r);
    float r2 = r * r;
    float ps = -1.9515295891e-4f;
    ps = ps * r2 + 8.3321608736e-3f;
    ps = ps * r2 + -1.6666654611e-1f;
    float sr = r + r * r2 * ps;
    float pc = 2.443315711809948e-5f;
    pc = pc * r2 + -1.388731625493765e-3f;
    pc = pc * r2 + 4.166664568298827e-2f;
    float cr = (1.0f - 0.5f * r2) + r2 * r2 * pc;
    switch (k & 3) {
    case 0: *sn = sr; *cs = cr; break;
    case 1: *sn = cr; *cs = -sr; break;
    case 2: *sn = -sr; *cs = -cr; break;
    default: *sn = -cr; *cs = sr; break;
    }
}
static inline rot rot_of(float a) { rot q; lite_sincosf(a, &q.s, &q.c); return q; }

/* ---------------------------------------------------------------- Box2D 2.3 constants (b2Settings.h) */
#define LINEAR_SLOP 0.005f
#define ANGULAR_SLOP (2.0f / 180.0f * 3.14159265359f)
#define POLYGON_RADIUS (2.0f * LINEAR_SLOP)
#define MAX_LINEAR_CORRECTION 0.2f
#define MAX_ANGULAR_CORRECTION (8.0f / 180.0f * 3.14159265359f)
#define MAX_TRANSLATION 2.0f
#define MAX_ROTATION (0.5f * 3.14159265359f)
#define BAUMGARTE 0.2f
#define VELOCITY_THRESHOLD 1.0f
#define TIME_TO_SLEEP 0.5f
#define LINEAR_SLEEP_TOL 0.01f
#define ANGULAR_SLEEP_TOL (2.0f / 180.0f * 3.14159265359f)
#define AABB_EXTENSION 0.1f

#define B2L_MAX_BODIES 5
#define B2L_MAX_CONTACTS 16
#define MAXV 6

typedef struct {
    int count;
    v2 verts[MAXV], normals[MAXV], centroid;
    float friction;
    /* body */
    xform xf;
    v2 localCenter, c0, c, v, force;
    float a0, a, w, torque, mass, invMass, I, invI, sleepTime;
    float alpha0;            /* b2Sweep::alpha0 (continuous collision, b2lite_toi.h) */
    int awake;
} body_t;

typedef struct { v2 v1, v2; float friction; } edge_t;

/* a static 4-gon (fixture of a static body at the identity transform): BipedalWalkerHardcore's stumps, stair steps
 * and pit walls.  verts / normals as b2PolygonShape::Set leaves them for an axis-aligned box. */
#define B2L_MAX_SPOLY 40
typedef struct { int count; v2 verts[4], normals[4]; float friction; } spoly_t;

typedef struct {
    v2 localPoint;
    float normalImpulse, tangentImpulse;
    uint32_t id;
} mpoint_t;

typedef struct {
    int type; /* 0 = faceA (edge is the reference face), 1 = faceB */
    v2 localNormal, localPoint;
    int pointCount;
    mpoint_t pts[2];
} manifold_t;

typedef struct { int touching; manifold_t m; } contact_t;

typedef struct {
    int bodyA, bodyB;
    v2 localAnchorA, localAnchorB;
    float lower, upper, maxMotorTorque, motorSpeed, referenceAngle;
    float impulse[3], motorImpulse;
    int limitState;       /* 0 inactive, 1 atLower, 2 atUpper, 3 equal */
    /* solver temp */
    v2 rA, rB, lcA, lcB;
    float mA, mB, iA, iB, K[3][3], motorMass; /* K[col][row] like b2Mat33 ex/ey/ez */
} joint_t;

/* A scene: nb dynamic bodies, nj revolute joints, ne static ground edges (all fixtures of static
 * bodies sitting at the identity transform).  ct is the [nb][ne] contact table. */
typedef struct {
    int nb, nj, ne;
    body_t *b;
    joint_t *j;
    const edge_t *e;
    contact_t *ct;
    int np, np_cap;          /* static polygons in use / row stride of ctp */
    const spoly_t *sp;
    contact_t *ctp;          /* [nb][np_cap] contact table of the (body, static polygon) pairs */
    const int *body_order;   /* island order of the bodies (contacts are built in this order) */
    const int *joint_order;  /* island order of the joints */
    float inv_dt0;
    float gravity_y;         /* b2World gravity = (0, gravity_y) */
    int awake;               /* out: 0 when the island went to sleep in this step */
    int stat_contacts, stat_pos_iters; /* out (workload statistics): touching contacts, position iterations run */
    int max_contacts;        /* manifold-table capacity of the scene (same number as the CUDA scene's kMaxVC; at most
                              * B2L_MAX_CONTACTS).  A pair that would be the (max_contacts+1)-th touching one of a step is
                              * treated as NOT touching -- no manifold, no BeginContact, EndContact if it was touching --
                              * identically here and on the device, and counted in `overflowed` */
    int overflowed;          /* out: number of pairs dropped that way in this step (tests assert it stays 0) */
    void (*event)(void *ctx, int body, int begin); /* Begin/EndContact listener */
    void *ctx;
    int toi;                 /* 1: b2World::SolveTOI after the discrete solve (m_continuousPhysics, Box2D's default) */
    int one_static_body;     /* 1: every static fixture belongs to ONE static body (LunarLander's moon); 0: one static
                              * body per fixture (BipedalWalker's terrain).  Matters for the static sweeps' alpha0 */
    int stat_toi_calls, stat_toi_events; /* out: b2TimeOfImpact evaluations / TOI sub-steps of this step */
} b2l_world;

/* ---------------------------------------------------------------- shapes & mass (b2PolygonShape) */
static void poly_set(body_t *b, const v2 *hull, int n)
{ /* vertices are given already in hull (CCW) order; normals + centroid as b2PolygonShape::Set */
    b->count = n;
    for (int i = 0; i < n; i++) b->verts[i] = hull[i];
    for (int i = 0; i < n; i++) {
        int i2 = i + 1 < n ? i + 1 : 0;
        v2 edge = sub(b->verts[i2], b->verts[i]);
        v2 nr = crs_vs(edge, 1.0f);
        float len = sqrtf(nr.x * nr.x + nr.y * nr.y);
        float inv = 1.0f / len;
        b->normals[i] = V(inv * nr.x, inv * nr.y);
    }
    /* ComputeCentroid */
    v2 c = V(0.0f, 0.0f), pRef = V(0.0f, 0.0f);
    float area = 0.0f;
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < n; i++) {
        v2 p1 = pRef, p2 = b->verts[i], p3 = i + 1 < n ? b->verts[i + 1] : b->verts[0];
        v2 e1 = sub(p2, p1), e2 = sub(p3, p1);
        float D = crs(e1, e2);
        float tri = 0.5f * D;
        area += tri;
        c = add(c, scl(tri * inv3, add(add(p1, p2), p3)));
    }
    b->centroid = scl(1.0f / area, c);
}

static void poly_mass(body_t *b, float density)
{ /* b2PolygonShape::ComputeMass + b2Body::ResetMassData (single fixture) */
    int n = b->count;
    v2 center = V(0.0f, 0.0f), s = V(0.0f, 0.0f);
    float area = 0.0f, I = 0.0f;
    for (int i = 0; i < n; i++) s = add(s, b->verts[i]);
    s = scl(1.0f / (float)n, s);
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < n; i++) {
        v2 e1 = sub(b->verts[i], s), e2 = i + 1 < n ? sub(b->verts[i + 1], s) : sub(b->verts[0], s);
        float D = crs(e1, e2), tri = 0.5f * D;
        area += tri;
        center = add(center, scl(tri * inv3, add(e1, e2)));
        float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
        float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2, inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
        I += (0.25f * inv3 * D) * (intx2 + inty2);
    }
    float mass = density * area;
    center = scl(1.0f / area, center);
    v2 mcenter = add(center, s);
    float mI = density * I;
    mI += mass * (dot(mcenter, mcenter) - dot(center, center));
    /* ResetMassData */
    b->mass = mass;
    b->invMass = 1.0f / mass;
    v2 lc = scl(b->invMass, scl(mass, mcenter));
    float bI = mI - mass * dot(lc, lc);
    b->I = bI;
    b->invI = 1.0f / bI;
    b->localCenter = lc;
}

static void body_place(body_t *b, v2 pos, float angle)
{
    b->xf.p = pos;
    b->xf.q = rot_of(angle);
    b->c = xmul(b->xf, b->localCenter);
    b->c0 = b->c;
    b->a = b->a0 = angle;
    b->v = V(0.0f, 0.0f);
    b->w = 0.0f;
    b->force = V(0.0f, 0.0f);
    b->torque = 0.0f;
    b->sleepTime = 0.0f;
    b->awake = 1;
}

static inline void body_sync_xf(body_t *b)
{ /* b2Body::SynchronizeTransform */
    b->xf.q = rot_of(b->a);
    b->xf.p = sub(b->c, rmul(b->xf.q, b->localCenter));
}

/* ---------------------------------------------------------------- b2CollideEdgeAndPolygon */
typedef struct { v2 v; uint32_t id; } clipv_t;
#define ID(ia, ib, ta, tb) ((uint32_t)(ia) | ((uint32_t)(ib) << 8) | ((uint32_t)(ta) << 16) | ((uint32_t)(tb) << 24))
enum { F_VERTEX = 0, F_FACE = 1 };

static int clip_segment(clipv_t out[2], const clipv_t in[2], v2 normal, float offset, int vertexIndexA)
{
    int n = 0;
    float d0 = dot(normal, in[0].v) - offset, d1 = dot(normal, in[1].v) - offset;
    if (d0 <= 0.0f) out[n++] = in[0];
    if (d1 <= 0.0f) out[n++] = in[1];
    if (d0 * d1 < 0.0f) {
        float interp = d0 / (d0 - d1);
        out[n].v = add(in[0].v, scl(interp, sub(in[1].v, in[0].v)));
        out[n].id = ID(vertexIndexA, (in[0].id >> 8) & 0xff, F_VERTEX, F_FACE);
        n++;
    }
    return n;
}

static void collide_edge_polygon(manifold_t *m, const edge_t *eA, const body_t *pB)
{ /* the edge's body (the moon) sits at the identity transform, so xf = xfB */
    const xform xf = pB->xf;
    const int count = pB->count;
    v2 centroidB = xmul(xf, pB->centroid);
    v2 v1 = eA->v1, v2_ = eA->v2;
    v2 edge1 = sub(v2_, v1);
    float len = sqrtf(edge1.x * edge1.x + edge1.y * edge1.y);
    float inv = 1.0f / len;
    edge1 = V(edge1.x * inv, edge1.y * inv);
    v2 normal1 = V(edge1.y, -edge1.x);
    float offset1 = dot(normal1, sub(centroidB, v1));
    int front = offset1 >= 0.0f;
    v2 normal, lower, upper;
    if (front) { normal = normal1; lower = neg(normal1); upper = neg(normal1); }
    else { normal = neg(normal1); lower = normal1; upper = normal1; }
    v2 pv[MAXV], pn[MAXV];
    for (int i = 0; i < count; i++) { pv[i] = xmul(xf, pB->verts[i]); pn[i] = rmul(xf.q, pB->normals[i]); }
    const float radius = 2.0f * POLYGON_RADIUS;
    m->pointCount = 0;
    /* ComputeEdgeSeparation */
    float edgeSep = 3.402823466e+38f;
    for (int i = 0; i < count; i++) { float s = dot(normal, sub(pv[i], v1)); if (s < edgeSep) edgeSep = s; }
    if (edgeSep > radius) return;
    /* ComputePolygonSeparation */
    int polyType = 0 /*unknown*/, polyIndex = -1;
    float polySep = -3.402823466e+38f;
    v2 perp = V(-normal.y, normal.x);
    for (int i = 0; i < count; i++) {
        v2 n = neg(pn[i]);
        float s1 = dot(n, sub(pv[i], v1)), s2 = dot(n, sub(pv[i], v2_));
        float s = fminf_(s1, s2);
        if (s > radius) { polyType = 1; polyIndex = i; polySep = s; break; }
        if (dot(n, perp) >= 0.0f) { if (dot(sub(n, upper), normal) < -ANGULAR_SLOP) continue; }
        else { if (dot(sub(n, lower), normal) < -ANGULAR_SLOP) continue; }
        if (s > polySep) { polyType = 1; polyIndex = i; polySep = s; }
    }
    if (polyType != 0 && polySep > radius) return;
    int primaryEdge; /* 1: edge A is the reference face */
    if (polyType == 0) primaryEdge = 1;
    else if (polySep > 0.98f * edgeSep + 0.001f) primaryEdge = 0;
    else primaryEdge = 1;

    clipv_t ie[2];
    int rf_i1, rf_i2;
    v2 rf_v1, rf_v2, rf_normal;
    if (primaryEdge) {
        m->type = 0;
        int best = 0;
        float bestVal = dot(normal, pn[0]);
        for (int i = 1; i < count; i++) { float val = dot(normal, pn[i]); if (val < bestVal) { bestVal = val; best = i; } }
        int i1 = best, i2 = i1 + 1 < count ? i1 + 1 : 0;
        ie[0].v = pv[i1]; ie[0].id = ID(0, i1, F_FACE, F_VERTEX);
        ie[1].v = pv[i2]; ie[1].id = ID(0, i2, F_FACE, F_VERTEX);
        if (front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = v1; rf_v2 = v2_; rf_normal = normal1; }
        else { rf_i1 = 1; rf_i2 = 0; rf_v1 = v2_; rf_v2 = v1; rf_normal = neg(normal1); }
    } else {
        m->type = 1;
        ie[0].v = v1; ie[0].id = ID(0, polyIndex, F_VERTEX, F_FACE);
        ie[1].v = v2_; ie[1].id = ID(0, polyIndex, F_VERTEX, F_FACE);
        rf_i1 = polyIndex; rf_i2 = rf_i1 + 1 < count ? rf_i1 + 1 : 0;
        rf_v1 = pv[rf_i1]; rf_v2 = pv[rf_i2]; rf_normal = pn[rf_i1];
    }
    v2 side1 = V(rf_normal.y, -rf_normal.x), side2 = neg(side1);
    float off1 = dot(side1, rf_v1), off2 = dot(side2, rf_v2);
    clipv_t c1[2], c2[2];
    if (clip_segment(c1, ie, side1, off1, rf_i1) < 2) return;
    if (clip_segment(c2, c1, side2, off2, rf_i2) < 2) return;
    if (primaryEdge) { m->localNormal = rf_normal; m->localPoint = rf_v1; }
    else { m->localNormal = pB->normals[rf_i1]; m->localPoint = pB->verts[rf_i1]; }
    int pc = 0;
    for (int i = 0; i < 2; i++) {
        float sep = dot(rf_normal, sub(c2[i].v, rf_v1));
        if (sep <= radius) {
            mpoint_t *cp = &m->pts[pc];
            if (primaryEdge) { cp->localPoint = xmulT(xf, c2[i].v); cp->id = c2[i].id; }
            else {
                uint32_t id = c2[i].id;
                cp->localPoint = c2[i].v;
                cp->id = ID((id >> 8) & 0xff, id & 0xff, (id >> 24) & 0xff, (id >> 16) & 0xff);
            }
            pc++;
        }
    }
    m->pointCount = pc;
}

/* b2Contact::Update for one (body, edge) pair; returns +1 on BeginContact, -1 on EndContact */
static int contact_update(contact_t *c, const edge_t *e, const body_t *b)
{
    manifold_t old = c->m;
    int was = c->touching;
    /* broad-phase stand-in: fat AABBs (b2_aabbExtension) must overlap, else no manifold */
    float lox = 3.402823466e+38f, loy = lox, hix = -lox, hiy = -lox;
    for (int i = 0; i < b->count; i++) {
        v2 p = xmul(b->xf, b->verts[i]);
        lox = fminf_(lox, p.x); loy = fminf_(loy, p.y); hix = fmaxf_(hix, p.x); hiy = fmaxf_(hiy, p.y);
    }
    const float ext = POLYGON_RADIUS + AABB_EXTENSION;
    float elox = fminf_(e->v1.x, e->v2.x) - ext, ehix = fmaxf_(e->v1.x, e->v2.x) + ext;
    float eloy = fminf_(e->v1.y, e->v2.y) - ext, ehiy = fmaxf_(e->v1.y, e->v2.y) + ext;
    c->m.pointCount = 0;
    if (!(lox - ext > ehix || elox > hix + ext || loy - ext > ehiy || eloy > hiy + ext))
        collide_edge_polygon(&c->m, e, b);
    int touching = c->m.pointCount > 0;
    for (int i = 0; i < c->m.pointCount; i++) {
        mpoint_t *mp2 = &c->m.pts[i];
        mp2->normalImpulse = 0.0f;
        mp2->tangentImpulse = 0.0f;
        if (was)
            for (int j = 0; j < old.pointCount; j++)
                if (old.pts[j].id == mp2->id) {
                    mp2->normalImpulse = old.pts[j].normalImpulse;
                    mp2->tangentImpulse = old.pts[j].tangentImpulse;
                    break;
                }
    }
    c->touching = touching;
    return touching - was;
}

/* ---------------------------------------------------------------- static polygons (b2CollidePolygons, v2.3.1+) */
/* b2PolygonShape::Set on the four corners of an axis-aligned box: the gift wrapping starts at the right-most
 * (lowest) corner and runs counter-clockwise; normals = normalised b2Cross(edge, 1). */
static void spoly_set_box(spoly_t *s, float x0, float ylo, float x1, float yhi, float friction)
{
    s->count = 4;
    s->verts[0] = V(x1, ylo); s->verts[1] = V(x1, yhi); s->verts[2] = V(x0, yhi); s->verts[3] = V(x0, ylo);
    for (int i = 0; i < 4; i++) {
        int i2 = i + 1 < 4 ? i + 1 : 0;
        v2 edge = sub(s->verts[i2], s->verts[i]);
        v2 nr = crs_vs(edge, 1.0f);
        float len = sqrtf(nr.x * nr.x + nr.y * nr.y);
        float inv = 1.0f / len;
        s->normals[i] = V(inv * nr.x, inv * nr.y);
    }
    s->friction = friction;
}

static inline rot rot_mulT(rot q, rot r) { rot o; o.s = q.c * r.s - q.s * r.c; o.c = q.c * r.c + q.s * r.s; return o; }
static inline xform xf_mulT(xform A, xform B) { xform C; C.q = rot_mulT(A.q, B.q); C.p = rmulT(A.q, sub(B.p, A.p)); return C; }

/* b2FindMaxSeparation(poly1, xf1, poly2, xf2) */
static float find_max_separation(int *edgeIndex, int count1, const v2 *n1s, const v2 *v1s, xform xf1, int count2,
                                 const v2 *v2s, xform xf2)
{
    xform xf = xf_mulT(xf2, xf1);
    int bestIndex = 0;
    float maxSeparation = -3.402823466e+38f;
    for (int i = 0; i < count1; i++) {
        v2 n = rmul(xf.q, n1s[i]);
        v2 v1 = xmul(xf, v1s[i]);
        float si = 3.402823466e+38f;
        for (int j = 0; j < count2; j++) { float sij = dot(n, sub(v2s[j], v1)); if (sij < si) si = sij; }
        if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
    }
    *edgeIndex = bestIndex;
    return maxSeparation;
}

/* b2CollidePolygons(manifold, polyA = static polygon at the identity, polyB = body polygon) */
static void collide_polygons(manifold_t *m, const spoly_t *A, const body_t *B)
{
    static const xform XFI = {{0.0f, 0.0f}, {0.0f, 1.0f}};
    const xform xfA = XFI, xfB = B->xf;
    m->pointCount = 0;
    const float totalRadius = POLYGON_RADIUS + POLYGON_RADIUS;
    int edgeA = 0, edgeB = 0;
    float separationA = find_max_separation(&edgeA, A->count, A->normals, A->verts, xfA, B->count, B->verts, xfB);
    if (separationA > totalRadius) return;
    float separationB = find_max_separation(&edgeB, B->count, B->normals, B->verts, xfB, A->count, A->verts, xfA);
    if (separationB > totalRadius) return;
    const v2 *verts1, *normals1, *verts2, *normals2;
    int count1, count2, edge1, flip;
    xform xf1, xf2;
    const float k_tol = 0.1f * LINEAR_SLOP;
    if (separationB > separationA + k_tol) {
        verts1 = B->verts; normals1 = B->normals; count1 = B->count; verts2 = A->verts; normals2 = A->normals; count2 = A->count;
        xf1 = xfB; xf2 = xfA; edge1 = edgeB; m->type = 1; flip = 1;
    } else {
        verts1 = A->verts; normals1 = A->normals; count1 = A->count; verts2 = B->verts; normals2 = B->normals; count2 = B->count;
        xf1 = xfA; xf2 = xfB; edge1 = edgeA; m->type = 0; flip = 0;
    }
    /* b2FindIncidentEdge */
    clipv_t ie[2];
    {
        v2 normal1 = rmulT(xf2.q, rmul(xf1.q, normals1[edge1]));
        int index = 0;
        float minDot = 3.402823466e+38f;
        for (int i = 0; i < count2; i++) { float d = dot(normal1, normals2[i]); if (d < minDot) { minDot = d; index = i; } }
        int i1 = index, i2 = i1 + 1 < count2 ? i1 + 1 : 0;
        ie[0].v = xmul(xf2, verts2[i1]); ie[0].id = ID(edge1, i1, F_FACE, F_VERTEX);
        ie[1].v = xmul(xf2, verts2[i2]); ie[1].id = ID(edge1, i2, F_FACE, F_VERTEX);
    }
    int iv1 = edge1, iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
    v2 v11 = verts1[iv1], v12 = verts1[iv2];
    v2 localTangent = sub(v12, v11);
    {
        float len = sqrtf(localTangent.x * localTangent.x + localTangent.y * localTangent.y);
        if (len >= 1.1920929e-07f) { float inv = 1.0f / len; localTangent.x *= inv; localTangent.y *= inv; }
    }
    v2 localNormal = crs_vs(localTangent, 1.0f);
    v2 planePoint = scl(0.5f, add(v11, v12));
    v2 tangent = rmul(xf1.q, localTangent);
    v2 normal = crs_vs(tangent, 1.0f);
    v11 = xmul(xf1, v11); v12 = xmul(xf1, v12);
    float frontOffset = dot(normal, v11);
    float sideOffset1 = -dot(tangent, v11) + totalRadius;
    float sideOffset2 = dot(tangent, v12) + totalRadius;
    clipv_t c1[2], c2[2];
    if (clip_segment(c1, ie, neg(tangent), sideOffset1, iv1) < 2) return;
    if (clip_segment(c2, c1, tangent, sideOffset2, iv2) < 2) return;
    m->localNormal = localNormal;
    m->localPoint = planePoint;
    int pc = 0;
    for (int i = 0; i < 2; i++) {
        float separation = dot(normal, c2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            mpoint_t *cp = &m->pts[pc];
            cp->localPoint = xmulT(xf2, c2[i].v);
            uint32_t id = c2[i].id;
            cp->id = flip ? ID((id >> 8) & 0xff, id & 0xff, (id >> 24) & 0xff, (id >> 16) & 0xff) : id;
            pc++;
        }
    }
    m->pointCount = pc;
}

/* b2Contact::Update for one (body, static polygon) pair; returns +1 on BeginContact, -1 on EndContact */
static int contact_update_poly(contact_t *c, const spoly_t *sp, const body_t *b)
{
    manifold_t old = c->m;
    int was = c->touching;
    float lox = 3.402823466e+38f, loy = lox, hix = -lox, hiy = -lox;
    for (int i = 0; i < b->count; i++) {
        v2 p = xmul(b->xf, b->verts[i]);
        lox = fminf_(lox, p.x); loy = fminf_(loy, p.y); hix = fmaxf_(hix, p.x); hiy = fmaxf_(hiy, p.y);
    }
    const float ext = POLYGON_RADIUS + AABB_EXTENSION;
    /* verts: 0 (x1, ylo), 2 (x0, yhi) */
    float elox = sp->verts[2].x - ext, ehix = sp->verts[0].x + ext, eloy = sp->verts[0].y - ext, ehiy = sp->verts[2].y + ext;
    c->m.pointCount = 0;
    if (!(lox - ext > ehix || elox > hix + ext || loy - ext > ehiy || eloy > hiy + ext))
        collide_polygons(&c->m, sp, b);
    int touching = c->m.pointCount > 0;
    for (int i = 0; i < c->m.pointCount; i++) {
        mpoint_t *mp2 = &c->m.pts[i];
        mp2->normalImpulse = 0.0f;
        mp2->tangentImpulse = 0.0f;
        if (was)
            for (int j = 0; j < old.pointCount; j++)
                if (old.pts[j].id == mp2->id) {
                    mp2->normalImpulse = old.pts[j].normalImpulse;
                    mp2->tangentImpulse = old.pts[j].tangentImpulse;
                    break;
                }
    }
    c->touching = touching;
    return touching - was;
}

/* b2PolygonShape::RayCast against a static polygon at the identity transform; returns 1 and *t on a hit */
static int spoly_raycast(const spoly_t *s, v2 p1_, v2 p2_, float maxFraction, float *t_out)
{
    static const xform XFI = {{0.0f, 0.0f}, {0.0f, 1.0f}};
    v2 p1 = rmulT(XFI.q, sub(p1_, XFI.p)), p2 = rmulT(XFI.q, sub(p2_, XFI.p));
    v2 d = sub(p2, p1);
    float lower = 0.0f, upper = maxFraction;
    int index = -1;
    for (int i = 0; i < s->count; i++) {
        float numerator = dot(s->normals[i], sub(s->verts[i], p1));
        float denominator = dot(s->normals[i], d);
        if (denominator == 0.0f) { if (numerator < 0.0f) return 0; }
        else {
            if (denominator < 0.0f && numerator < lower * denominator) { lower = numerator / denominator; index = i; }
            else if (denominator > 0.0f && numerator < upper * denominator) upper = numerator / denominator;
        }
        if (upper < lower) return 0;
    }
    if (index >= 0) { *t_out = lower; return 1; }
    return 0;
}

/* ---------------------------------------------------------------- contact solver (b2ContactSolver) */
typedef struct { v2 rA, rB; float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias; } vcp_t;
typedef struct {
    int body, edge, pointCount;
    v2 normal;
    vcp_t p[2];
    float K[2][2], nM[2][2]; /* [col][row] */
    float friction;
    manifold_t *m;
} vc_t;

typedef struct { v2 c; float a; v2 v; float w; } bstate_t;

static void world_manifold(const manifold_t *m, xform xfA, xform xfB, float rA, float rB, v2 *normal, v2 pts[2])
{ /* b2WorldManifold::Initialize for e_faceA / e_faceB */
    if (m->type == 0) {
        *normal = rmul(xfA.q, m->localNormal);
        v2 plane = xmul(xfA, m->localPoint);
        for (int i = 0; i < m->pointCount; i++) {
            v2 clip = xmul(xfB, m->pts[i].localPoint);
            v2 cA = add(clip, scl(rA - dot(sub(clip, plane), *normal), *normal));
            v2 cB = sub(clip, scl(rB, *normal));
            pts[i] = scl(0.5f, add(cA, cB));
        }
    } else {
        *normal = rmul(xfB.q, m->localNormal);
        v2 plane = xmul(xfB, m->localPoint);
        for (int i = 0; i < m->pointCount; i++) {
            v2 clip = xmul(xfA, m->pts[i].localPoint);
            v2 cB = add(clip, scl(rB - dot(sub(clip, plane), *normal), *normal));
            v2 cA = sub(clip, scl(rA, *normal));
            pts[i] = scl(0.5f, add(cA, cB));
        }
        *normal = neg(*normal);
    }
}

static const xform XF_ID = {{0.0f, 0.0f}, {0.0f, 1.0f}};

/* ---------------------------------------------------------------- b2RevoluteJoint */
static void mat33_solve33(float K[3][3], const float b[3], float x[3])
{
    const float *ex = K[0], *ey = K[1], *ez = K[2];
    float cx = ey[1] * ez[2] - ey[2] * ez[1], cy = ey[2] * ez[0] - ey[0] * ez[2], cz = ey[0] * ez[1] - ey[1] * ez[0];
    float det = ex[0] * cx + ex[1] * cy + ex[2] * cz;
    if (det != 0.0f) det = 1.0f / det;
    x[0] = det * (b[0] * cx + b[1] * cy + b[2] * cz);
    float bx = b[1] * ez[2] - b[2] * ez[1], by = b[2] * ez[0] - b[0] * ez[2], bz = b[0] * ez[1] - b[1] * ez[0];
    x[1] = det * (ex[0] * bx + ex[1] * by + ex[2] * bz);
    float dx = ey[1] * b[2] - ey[2] * b[1], dy = ey[2] * b[0] - ey[0] * b[2], dz = ey[0] * b[1] - ey[1] * b[0];
    x[2] = det * (ex[0] * dx + ex[1] * dy + ex[2] * dz);
}
static v2 mat33_solve22(float K[3][3], v2 b)
{
    float a11 = K[0][0], a12 = K[1][0], a21 = K[0][1], a22 = K[1][1];
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    return V(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
}

static void joint_init(joint_t *j, const body_t *A, const body_t *B, bstate_t *sA, bstate_t *sB, float dtRatio)
{
    j->lcA = A->localCenter; j->lcB = B->localCenter;
    j->mA = A->invMass; j->mB = B->invMass; j->iA = A->invI; j->iB = B->invI;
    rot qA = rot_of(sA->a), qB = rot_of(sB->a);
    j->rA = rmul(qA, sub(j->localAnchorA, j->lcA));
    j->rB = rmul(qB, sub(j->localAnchorB, j->lcB));
    float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
    v2 rA = j->rA, rB = j->rB;
    j->K[0][0] = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
    j->K[1][0] = -rA.y * rA.x * iA - rB.y * rB.x * iB;
    j->K[2][0] = -rA.y * iA - rB.y * iB;
    j->K[0][1] = j->K[1][0];
    j->K[1][1] = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
    j->K[2][1] = rA.x * iA + rB.x * iB;
    j->K[0][2] = j->K[2][0];
    j->K[1][2] = j->K[2][1];
    j->K[2][2] = iA + iB;
    j->motorMass = iA + iB;
    if (j->motorMass > 0.0f) j->motorMass = 1.0f / j->motorMass;
    float jointAngle = sB->a - sA->a - j->referenceAngle;
    if (fabsf(j->upper - j->lower) < 2.0f * ANGULAR_SLOP) j->limitState = 3;
    else if (jointAngle <= j->lower) { if (j->limitState != 1) j->impulse[2] = 0.0f; j->limitState = 1; }
    else if (jointAngle >= j->upper) { if (j->limitState != 2) j->impulse[2] = 0.0f; j->limitState = 2; }
    else { j->limitState = 0; j->impulse[2] = 0.0f; }
    /* warm start */
    j->impulse[0] *= dtRatio; j->impulse[1] *= dtRatio; j->impulse[2] *= dtRatio; j->motorImpulse *= dtRatio;
    v2 P = V(j->impulse[0], j->impulse[1]);
    sA->v = sub(sA->v, scl(mA, P));
    sA->w -= iA * (crs(rA, P) + j->motorImpulse + j->impulse[2]);
    sB->v = add(sB->v, scl(mB, P));
    sB->w += iB * (crs(rB, P) + j->motorImpulse + j->impulse[2]);
}

static void joint_solve_velocity(joint_t *j, bstate_t *sA, bstate_t *sB, float dt)
{
    float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
    v2 vA = sA->v, vB = sB->v;
    float wA = sA->w, wB = sB->w;
    if (j->limitState != 3) { /* motor */
        float Cdot = wB - wA - j->motorSpeed;
        float impulse = -j->motorMass * Cdot;
        float oldImpulse = j->motorImpulse, maxImpulse = dt * j->maxMotorTorque;
        j->motorImpulse = clampf(oldImpulse + impulse, -maxImpulse, maxImpulse);
        impulse = j->motorImpulse - oldImpulse;
        wA -= iA * impulse;
        wB += iB * impulse;
    }
    if (j->limitState != 0) {
        v2 Cdot1 = sub(sub(add(vB, crs_sv(wB, j->rB)), vA), crs_sv(wA, j->rA));
        float Cdot2 = wB - wA;
        float Cd[3] = {Cdot1.x, Cdot1.y, Cdot2}, imp[3];
        mat33_solve33(j->K, Cd, imp);
        imp[0] = -imp[0]; imp[1] = -imp[1]; imp[2] = -imp[2];
        if (j->limitState == 3) { j->impulse[0] += imp[0]; j->impulse[1] += imp[1]; j->impulse[2] += imp[2]; }
        else {
            float newImpulse = j->impulse[2] + imp[2];
            int violated = (j->limitState == 1) ? (newImpulse < 0.0f) : (newImpulse > 0.0f);
            if (violated) {
                v2 rhs = add(neg(Cdot1), scl(j->impulse[2], V(j->K[2][0], j->K[2][1])));
                v2 red = mat33_solve22(j->K, rhs);
                imp[0] = red.x; imp[1] = red.y; imp[2] = -j->impulse[2];
                j->impulse[0] += red.x; j->impulse[1] += red.y; j->impulse[2] = 0.0f;
            } else { j->impulse[0] += imp[0]; j->impulse[1] += imp[1]; j->impulse[2] += imp[2]; }
        }
        v2 P = V(imp[0], imp[1]);
        vA = sub(vA, scl(mA, P)); wA -= iA * (crs(j->rA, P) + imp[2]);
        vB = add(vB, scl(mB, P)); wB += iB * (crs(j->rB, P) + imp[2]);
    } else {
        v2 Cdot = sub(sub(add(vB, crs_sv(wB, j->rB)), vA), crs_sv(wA, j->rA));
        v2 imp = mat33_solve22(j->K, neg(Cdot));
        j->impulse[0] += imp.x; j->impulse[1] += imp.y;
        vA = sub(vA, scl(mA, imp)); wA -= iA * crs(j->rA, imp);
        vB = add(vB, scl(mB, imp)); wB += iB * crs(j->rB, imp);
    }
    sA->v = vA; sA->w = wA; sB->v = vB; sB->w = wB;
}

static int joint_solve_position(joint_t *j, bstate_t *sA, bstate_t *sB)
{
    v2 cA = sA->c, cB = sB->c;
    float aA = sA->a, aB = sB->a, angularError = 0.0f, positionError;
    if (j->limitState != 0) {
        float angle = aB - aA - j->referenceAngle, limitImpulse = 0.0f;
        if (j->limitState == 3) {
            float C = clampf(angle - j->lower, -MAX_ANGULAR_CORRECTION, MAX_ANGULAR_CORRECTION);
            limitImpulse = -j->motorMass * C; angularError = fabsf(C);
        } else if (j->limitState == 1) {
            float C = angle - j->lower; angularError = -C;
            C = clampf(C + ANGULAR_SLOP, -MAX_ANGULAR_CORRECTION, 0.0f); limitImpulse = -j->motorMass * C;
        } else {
            float C = angle - j->upper; angularError = C;
            C = clampf(C - ANGULAR_SLOP, 0.0f, MAX_ANGULAR_CORRECTION); limitImpulse = -j->motorMass * C;
        }
        aA -= j->iA * limitImpulse;
        aB += j->iB * limitImpulse;
    }
    {
        rot qA = rot_of(aA), qB = rot_of(aB);
        v2 rA = rmul(qA, sub(j->localAnchorA, j->lcA)), rB = rmul(qB, sub(j->localAnchorB, j->lcB));
        v2 C = sub(sub(add(cB, rB), cA), rA);
        positionError = sqrtf(C.x * C.x + C.y * C.y);
        float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
        float k11 = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
        float k12 = -iA * rA.x * rA.y - iB * rB.x * rB.y;
        float k22 = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
        float det = k11 * k22 - k12 * k12;
        if (det != 0.0f) det = 1.0f / det;
        v2 imp = V(-(det * (k22 * C.x - k12 * C.y)), -(det * (k11 * C.y - k12 * C.x)));
        cA = sub(cA, scl(mA, imp)); aA -= iA * crs(rA, imp);
        cB = add(cB, scl(mB, imp)); aB += iB * crs(rB, imp);
    }
    sA->c = cA; sA->a = aA; sB->c = cB; sB->a = aB;
    return positionError <= LINEAR_SLOP && angularError <= ANGULAR_SLOP;
}

#include "b2lite_toi.h"

/* ---------------------------------------------------------------- b2World::Step */
static void b2l_step(b2l_world *W, float dt, int velIters, int posIters)
{
    const float inv_dt = 1.0f / dt;
    const float dtRatio = W->inv_dt0 * dt;
    const v2 gravity = V(0.0f, W->gravity_y);
    /* --- Collide: update manifolds, begin/end events (the env's ContactDetector);
     *     pairs are visited in the same (island) order the constraints are built in */
    const int NB = W->nb, NE = W->ne;
    const int *order = W->body_order, *jorder = W->joint_order;
    const int cap = (W->max_contacts > 0 && W->max_contacts < B2L_MAX_CONTACTS) ? W->max_contacts : B2L_MAX_CONTACTS;
    int ntouch = 0;
    W->overflowed = 0;
    for (int oi = 0; oi < NB; oi++) {
        int b = order[oi];
        for (int p = W->np - 1; p >= 0; p--) {  /* static polygons first (descending), then the edges */
            contact_t *c = &W->ctp[b * W->np_cap + p];
            int was = c->touching;
            int ev = contact_update_poly(c, &W->sp[p], &W->b[b]);
            if (c->touching && ntouch >= cap) { c->touching = 0; c->m.pointCount = 0; W->overflowed++; ev = -was; }
            ntouch += c->touching;
            if (ev != 0 && W->event) W->event(W->ctx, b, ev > 0);
        }
        for (int e = NE - 1; e >= 0; e--) {
            contact_t *c = &W->ct[b * NE + e];
            int was = c->touching;
            int ev = contact_update(c, &W->e[e], &W->b[b]);
            if (c->touching && ntouch >= cap) { c->touching = 0; c->m.pointCount = 0; W->overflowed++; ev = -was; }
            ntouch += c->touching;
            if (ev != 0 && W->event) W->event(W->ctx, b, ev > 0);
        }
    }
    /* --- Solve: one island {leg+1, lander, leg-1} (+ the static moon) */
    bstate_t st[B2L_MAX_BODIES];
    for (int i = 0; i < NB; i++) {
        body_t *b = &W->b[i];
        b->c0 = b->c; b->a0 = b->a;
        v2 v = b->v; float w = b->w;
        v = add(v, scl(dt, add(gravity, scl(b->invMass, b->force))));
        w = w + dt * b->invI * b->torque;
        v = scl(1.0f / (1.0f + dt * 0.0f), v);
        w = w * (1.0f / (1.0f + dt * 0.0f));
        st[i].c = b->c; st[i].a = b->a; st[i].v = v; st[i].w = w;
    }
    /* contact constraints in island order */
    vc_t vc[B2L_MAX_CONTACTS];
    int nvc = 0;
    for (int oi = 0; oi < NB; oi++) {
        int b = order[oi];
        for (int f = 0; f < W->np + NE; f++) {
            /* f < np: polygon np-1-f; else edge NE-1-(f-np) */
            const int isp = f < W->np;
            const int e = isp ? W->np - 1 - f : NE - 1 - (f - W->np);
            contact_t *c = isp ? &W->ctp[b * W->np_cap + e] : &W->ct[b * NE + e];
            if (!c->touching || nvc >= B2L_MAX_CONTACTS) continue;
            vc_t *k = &vc[nvc++];
            k->body = b; k->edge = isp ? NE + e : e; k->m = &c->m; k->pointCount = c->m.pointCount;
            k->friction = sqrtf((isp ? W->sp[e].friction : W->e[e].friction) * W->b[b].friction);
            for (int p = 0; p < k->pointCount; p++) {
                k->p[p].normalImpulse = dtRatio * c->m.pts[p].normalImpulse;
                k->p[p].tangentImpulse = dtRatio * c->m.pts[p].tangentImpulse;
            }
        }
    }
    /* InitializeVelocityConstraints */
    for (int ci = 0; ci < nvc; ci++) {
        vc_t *k = &vc[ci];
        body_t *B = &W->b[k->body];
        bstate_t *sB = &st[k->body];
        float mB = B->invMass, iB = B->invI;
        xform xfB;
        xfB.q = rot_of(sB->a);
        xfB.p = sub(sB->c, rmul(xfB.q, B->localCenter));
        v2 pts[2];
        world_manifold(k->m, XF_ID, xfB, POLYGON_RADIUS, POLYGON_RADIUS, &k->normal, pts);
        for (int p = 0; p < k->pointCount; p++) {
            vcp_t *cp = &k->p[p];
            cp->rA = pts[p]; /* cA = 0 */
            cp->rB = sub(pts[p], sB->c);
            float rnB = crs(cp->rB, k->normal);
            float kN = mB + iB * rnB * rnB;
            cp->normalMass = kN > 0.0f ? 1.0f / kN : 0.0f;
            v2 tangent = crs_vs(k->normal, 1.0f);
            float rtB = crs(cp->rB, tangent);
            float kT = mB + iB * rtB * rtB;
            cp->tangentMass = kT > 0.0f ? 1.0f / kT : 0.0f;
            cp->velocityBias = 0.0f;
            float vRel = dot(k->normal, add(sB->v, crs_sv(sB->w, cp->rB)));
            if (vRel < -VELOCITY_THRESHOLD) cp->velocityBias = -0.0f * vRel;
        }
        if (k->pointCount == 2) {
            float rn1B = crs(k->p[0].rB, k->normal), rn2B = crs(k->p[1].rB, k->normal);
            float k11 = mB + iB * rn1B * rn1B, k22 = mB + iB * rn2B * rn2B, k12 = mB + iB * rn1B * rn2B;
            if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
                k->K[0][0] = k11; k->K[0][1] = k12; k->K[1][0] = k12; k->K[1][1] = k22;
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = 1.0f / det;
                k->nM[0][0] = det * k22; k->nM[1][0] = -det * k12; k->nM[0][1] = -det * k12; k->nM[1][1] = det * k11;
            } else k->pointCount = 1;
        }
    }
    /* WarmStart */
    for (int ci = 0; ci < nvc; ci++) {
        vc_t *k = &vc[ci];
        body_t *B = &W->b[k->body];
        bstate_t *sB = &st[k->body];
        v2 tangent = crs_vs(k->normal, 1.0f);
        for (int p = 0; p < k->pointCount; p++) {
            v2 P = add(scl(k->p[p].normalImpulse, k->normal), scl(k->p[p].tangentImpulse, tangent));
            sB->w += B->invI * crs(k->p[p].rB, P);
            sB->v = add(sB->v, scl(B->invMass, P));
        }
    }
    /* joints in island order */
    const int NJ = W->nj;
    for (int q = 0; q < NJ; q++) { joint_t *j = &W->j[jorder[q]]; joint_init(j, &W->b[j->bodyA], &W->b[j->bodyB], &st[j->bodyA], &st[j->bodyB], dtRatio); }
    /* velocity iterations */
    for (int it = 0; it < velIters; it++) {
        for (int q = 0; q < NJ; q++) { joint_t *j = &W->j[jorder[q]]; joint_solve_velocity(j, &st[j->bodyA], &st[j->bodyB], dt); }
        for (int ci = 0; ci < nvc; ci++) {
            vc_t *k = &vc[ci];
            body_t *B = &W->b[k->body];
            float mB = B->invMass, iB = B->invI;
            v2 vB = st[k->body].v; float wB = st[k->body].w;
            v2 normal = k->normal, tangent = crs_vs(normal, 1.0f);
            for (int p = 0; p < k->pointCount; p++) {
                vcp_t *cp = &k->p[p];
                v2 dv = add(vB, crs_sv(wB, cp->rB));
                float vt = dot(dv, tangent) - 0.0f;
                float lambda = cp->tangentMass * (-vt);
                float maxF = k->friction * cp->normalImpulse;
                float newImp = clampf(cp->tangentImpulse + lambda, -maxF, maxF);
                lambda = newImp - cp->tangentImpulse;
                cp->tangentImpulse = newImp;
                v2 P = scl(lambda, tangent);
                vB = add(vB, scl(mB, P)); wB += iB * crs(cp->rB, P);
            }
            if (k->pointCount == 1) {
                vcp_t *cp = &k->p[0];
                v2 dv = add(vB, crs_sv(wB, cp->rB));
                float vn = dot(dv, normal);
                float lambda = -cp->normalMass * (vn - cp->velocityBias);
                float newImp = fmaxf_(cp->normalImpulse + lambda, 0.0f);
                lambda = newImp - cp->normalImpulse;
                cp->normalImpulse = newImp;
                v2 P = scl(lambda, normal);
                vB = add(vB, scl(mB, P)); wB += iB * crs(cp->rB, P);
            } else {
                vcp_t *c1 = &k->p[0], *c2 = &k->p[1];
                v2 a = V(c1->normalImpulse, c2->normalImpulse);
                v2 dv1 = add(vB, crs_sv(wB, c1->rB)), dv2 = add(vB, crs_sv(wB, c2->rB));
                float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
                v2 b = V(vn1 - c1->velocityBias, vn2 - c2->velocityBias);
                b = sub(b, V(k->K[0][0] * a.x + k->K[1][0] * a.y, k->K[0][1] * a.x + k->K[1][1] * a.y));
                v2 x;
                int solved = 0;
                x = V(-(k->nM[0][0] * b.x + k->nM[1][0] * b.y), -(k->nM[0][1] * b.x + k->nM[1][1] * b.y));
                if (x.x >= 0.0f && x.y >= 0.0f) solved = 1;
                if (!solved) {
                    x.x = -c1->normalMass * b.x; x.y = 0.0f;
                    vn2 = k->K[0][1] * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) solved = 1;
                }
                if (!solved) {
                    x.x = 0.0f; x.y = -c2->normalMass * b.y;
                    vn1 = k->K[1][0] * x.y + b.x;
                    if (x.y >= 0.0f && vn1 >= 0.0f) solved = 1;
                }
                if (!solved) {
                    x.x = 0.0f; x.y = 0.0f;
                    if (b.x >= 0.0f && b.y >= 0.0f) solved = 1;
                }
                if (solved) {
                    v2 d = sub(x, a);
                    v2 P1 = scl(d.x, normal), P2 = scl(d.y, normal);
                    vB = add(vB, scl(mB, add(P1, P2)));
                    wB += iB * (crs(c1->rB, P1) + crs(c2->rB, P2));
                    c1->normalImpulse = x.x; c2->normalImpulse = x.y;
                }
            }
            st[k->body].v = vB; st[k->body].w = wB;
        }
    }
    /* StoreImpulses */
    for (int ci = 0; ci < nvc; ci++)
        for (int p = 0; p < vc[ci].pointCount; p++) {
            vc[ci].m->pts[p].normalImpulse = vc[ci].p[p].normalImpulse;
            vc[ci].m->pts[p].tangentImpulse = vc[ci].p[p].tangentImpulse;
        }
    /* integrate positions */
    for (int i = 0; i < NB; i++) {
        v2 v = st[i].v; float w = st[i].w;
        v2 tr = scl(dt, v);
        if (dot(tr, tr) > MAX_TRANSLATION * MAX_TRANSLATION) { float ratio = MAX_TRANSLATION / sqrtf(dot(tr, tr)); v = scl(ratio, v); }
        float rotn = dt * w;
        if (rotn * rotn > MAX_ROTATION * MAX_ROTATION) { float ratio = MAX_ROTATION / fabsf(rotn); w *= ratio; }
        st[i].c = add(st[i].c, scl(dt, v));
        st[i].a = st[i].a + dt * w;
        st[i].v = v; st[i].w = w;
    }
    /* position iterations */
    int positionSolved = 0;
    W->stat_contacts = nvc; W->stat_pos_iters = 0;
    for (int it = 0; it < posIters; it++) {
        W->stat_pos_iters = it + 1;
        float minSep = 0.0f;
        for (int ci = 0; ci < nvc; ci++) {
            vc_t *k = &vc[ci];
            body_t *B = &W->b[k->body];
            float mB = B->invMass, iB = B->invI;
            v2 cB = st[k->body].c; float aB = st[k->body].a;
            int npts = k->m->pointCount; /* the position solver uses the manifold's own point count */
            for (int p = 0; p < npts; p++) {
                xform xfB;
                xfB.q = rot_of(aB);
                xfB.p = sub(cB, rmul(xfB.q, B->localCenter));
                v2 normal, point;
                float separation;
                if (k->m->type == 0) {
                    normal = k->m->localNormal;
                    v2 plane = k->m->localPoint;
                    v2 clip = xmul(xfB, k->m->pts[p].localPoint);
                    separation = dot(sub(clip, plane), normal) - POLYGON_RADIUS - POLYGON_RADIUS;
                    point = clip;
                } else {
                    normal = rmul(xfB.q, k->m->localNormal);
                    v2 plane = xmul(xfB, k->m->localPoint);
                    v2 clip = k->m->pts[p].localPoint;
                    separation = dot(sub(clip, plane), normal) - POLYGON_RADIUS - POLYGON_RADIUS;
                    point = clip;
                    normal = neg(normal);
                }
                v2 rB = sub(point, cB);
                minSep = fminf_(minSep, separation);
                float C = clampf(BAUMGARTE * (separation + LINEAR_SLOP), -MAX_LINEAR_CORRECTION, 0.0f);
                float rnB = crs(rB, normal);
                float K = mB + iB * rnB * rnB;
                float impulse = K > 0.0f ? -C / K : 0.0f;
                v2 P = scl(impulse, normal);
                cB = add(cB, scl(mB, P));
                aB += iB * crs(rB, P);
            }
            st[k->body].c = cB; st[k->body].a = aB;
        }
        int contactsOkay = minSep >= -3.0f * LINEAR_SLOP;
        int jointsOkay = 1;
        for (int q = 0; q < NJ; q++) {
            joint_t *j = &W->j[jorder[q]];
            int ok = joint_solve_position(j, &st[j->bodyA], &st[j->bodyB]);
            jointsOkay = jointsOkay && ok;
        }
        if (contactsOkay && jointsOkay) { positionSolved = 1; break; }
    }
    /* copy back, sleep */
    for (int i = 0; i < NB; i++) {
        body_t *b = &W->b[i];
        b->c = st[i].c; b->a = st[i].a; b->v = st[i].v; b->w = st[i].w;
        body_sync_xf(b);
    }
    float minSleep = 3.402823466e+38f;
    const float linTol = LINEAR_SLEEP_TOL * LINEAR_SLEEP_TOL, angTol = ANGULAR_SLEEP_TOL * ANGULAR_SLEEP_TOL;
    for (int i = 0; i < NB; i++) {
        body_t *b = &W->b[i];
        if (b->w * b->w > angTol || dot(b->v, b->v) > linTol) { b->sleepTime = 0.0f; minSleep = 0.0f; }
        else { b->sleepTime += dt; minSleep = fminf_(minSleep, b->sleepTime); }
    }
    W->awake = 1;
    if (minSleep >= TIME_TO_SLEEP && positionSolved) {
        W->awake = 0;
        for (int i = 0; i < NB; i++) {
            body_t *b = &W->b[i];
            b->awake = 0; b->sleepTime = 0.0f; b->v = V(0.0f, 0.0f); b->w = 0.0f;
        }
    }
    /* --- SolveTOI: continuous collision against the static fixtures (sleeping islands are skipped) */
    W->stat_toi_calls = 0; W->stat_toi_events = 0;
    if (W->toi && W->awake) b2l_solve_toi(W, dt, velIters);
    for (int i = 0; i < NB; i++) { W->b[i].force = V(0.0f, 0.0f); W->b[i].torque = 0.0f; }
    W->inv_dt0 = inv_dt;
}


#endif
