/*
 * b2lite_toi.h -- CPU ORACLE building block (test infrastructure, NOT product code); included by b2lite.h.
 *
 * Continuous collision of Box2D 2.3: b2World::SolveTOI with b2TimeOfImpact (conservative advancement on a
 * separating axis, b2SeparationFunction + b2Distance/GJK with a simplex cache) and b2Island::SolveTOI
 * (b2ContactSolver::SolveTOIPositionConstraints with b2_toiBaugarte = 0.75, 20 position iterations, the step's
 * velocity iterations without warm starting, integration over the rest of the step), restated from the published
 * algorithm for the only case gym's Box2D tasks exercise: a non-bullet dynamic polygon against fixtures of static
 * bodies (edges, and BipedalWalkerHardcore's boxes).  Dynamic-vs-dynamic pairs never get a TOI in Box2D unless one
 * is a bullet, and the tasks' dynamic bodies do not collide with each other at all (category / mask bits).
 * PARITY UNPINNED like the rest of b2lite.h.  Stand-ins for the broad phase: the contact list of SolveTOI is
 * "every (body, static fixture) pair whose fat AABBs overlap, the body's AABB taken over its sweep" (what
 * b2Body::SynchronizeFixtures hands to the dynamic tree, without the tree's displacement prediction / hysteresis:
 * the extra pairs those would add are separated over the whole step and cannot produce an event), ordered like
 * every other per-pair loop here (island order of the bodies x descending fixture index).
 */
#ifndef B2LITE_TOI_H
#define B2LITE_TOI_H

#define B2_EPSILON 1.1920929e-07f
#define B2_PI 3.14159265359f
#define MAX_SUB_STEPS 8
#define TOI_BAUMGARTE 0.75f
#define B2L_MAX_TOI_CAND 48

/* ---------------------------------------------------------------- b2Sweep */
typedef struct { v2 localCenter, c0, c; float a0, a, alpha0; } sweep_t;

static xform sweep_xf(const sweep_t *s, float beta)
{ /* b2Sweep::GetTransform */
    xform xf;
    xf.p = add(scl(1.0f - beta, s->c0), scl(beta, s->c));
    float angle = (1.0f - beta) * s->a0 + beta * s->a;
    xf.q = rot_of(angle);
    xf.p = sub(xf.p, rmul(xf.q, s->localCenter));
    return xf;
}
static void sweep_advance(sweep_t *s, float alpha)
{ /* b2Sweep::Advance */
    float beta = (alpha - s->alpha0) / (1.0f - s->alpha0);
    s->c0 = add(s->c0, scl(beta, sub(s->c, s->c0)));
    s->a0 = s->a0 + beta * (s->a - s->a0);
    s->alpha0 = alpha;
}
static void sweep_normalize(sweep_t *s)
{ /* b2Sweep::Normalize */
    float twoPi = 2.0f * B2_PI;
    float d = twoPi * floorf(s->a0 / twoPi);
    s->a0 -= d;
    s->a -= d;
}

/* ---------------------------------------------------------------- b2Distance (GJK) */
typedef struct { int count; v2 v[MAXV]; } dproxy_t;
typedef struct { v2 wA, wB, w; float a; int indexA, indexB; } sv_t;
typedef struct { sv_t v[3]; int count; } simplex_t;
typedef struct { float metric; int count; int indexA[3], indexB[3]; } scache_t;

static int proxy_support(const dproxy_t *p, v2 d)
{
    int best = 0;
    float bestValue = dot(p->v[0], d);
    for (int i = 1; i < p->count; i++) { float value = dot(p->v[i], d); if (value > bestValue) { best = i; bestValue = value; } }
    return best;
}
static float dist2(v2 a, v2 b) { v2 c = sub(a, b); return sqrtf(c.x * c.x + c.y * c.y); }

static float simplex_metric(const simplex_t *s)
{
    if (s->count == 2) return dist2(s->v[0].w, s->v[1].w);
    if (s->count == 3) return crs(sub(s->v[1].w, s->v[0].w), sub(s->v[2].w, s->v[0].w));
    return 0.0f;
}
static void simplex_read_cache(simplex_t *s, const scache_t *cache, const dproxy_t *pA, xform xfA, const dproxy_t *pB, xform xfB)
{
    s->count = cache->count;
    for (int i = 0; i < s->count; i++) {
        sv_t *v = &s->v[i];
        v->indexA = cache->indexA[i]; v->indexB = cache->indexB[i];
        v->wA = xmul(xfA, pA->v[v->indexA]); v->wB = xmul(xfB, pB->v[v->indexB]);
        v->w = sub(v->wB, v->wA); v->a = 0.0f;
    }
    if (s->count > 1) {
        float metric1 = cache->metric, metric2 = simplex_metric(s);
        if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < B2_EPSILON) s->count = 0;
    }
    if (s->count == 0) {
        sv_t *v = &s->v[0];
        v->indexA = 0; v->indexB = 0;
        v->wA = xmul(xfA, pA->v[0]); v->wB = xmul(xfB, pB->v[0]);
        v->w = sub(v->wB, v->wA); v->a = 1.0f;
        s->count = 1;
    }
}
static void simplex_write_cache(const simplex_t *s, scache_t *cache)
{
    cache->metric = simplex_metric(s);
    cache->count = s->count;
    for (int i = 0; i < s->count; i++) { cache->indexA[i] = s->v[i].indexA; cache->indexB[i] = s->v[i].indexB; }
}
static void simplex_solve2(simplex_t *s)
{
    v2 w1 = s->v[0].w, w2 = s->v[1].w, e12 = sub(w2, w1);
    float d12_2 = -dot(w1, e12);
    if (d12_2 <= 0.0f) { s->v[0].a = 1.0f; s->count = 1; return; }
    float d12_1 = dot(w2, e12);
    if (d12_1 <= 0.0f) { s->v[1].a = 1.0f; s->count = 1; s->v[0] = s->v[1]; return; }
    float inv_d12 = 1.0f / (d12_1 + d12_2);
    s->v[0].a = d12_1 * inv_d12; s->v[1].a = d12_2 * inv_d12; s->count = 2;
}
static void simplex_solve3(simplex_t *s)
{
    v2 w1 = s->v[0].w, w2 = s->v[1].w, w3 = s->v[2].w;
    v2 e12 = sub(w2, w1);
    float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12), d12_1 = w2e12, d12_2 = -w1e12;
    v2 e13 = sub(w3, w1);
    float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13), d13_1 = w3e13, d13_2 = -w1e13;
    v2 e23 = sub(w3, w2);
    float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23), d23_1 = w3e23, d23_2 = -w2e23;
    float n123 = crs(e12, e13);
    float d123_1 = n123 * crs(w2, w3), d123_2 = n123 * crs(w3, w1), d123_3 = n123 * crs(w1, w2);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) { s->v[0].a = 1.0f; s->count = 1; return; }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
        float inv = 1.0f / (d12_1 + d12_2);
        s->v[0].a = d12_1 * inv; s->v[1].a = d12_2 * inv; s->count = 2; return;
    }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
        float inv = 1.0f / (d13_1 + d13_2);
        s->v[0].a = d13_1 * inv; s->v[2].a = d13_2 * inv; s->count = 2; s->v[1] = s->v[2]; return;
    }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) { s->v[1].a = 1.0f; s->count = 1; s->v[0] = s->v[1]; return; }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) { s->v[2].a = 1.0f; s->count = 1; s->v[0] = s->v[2]; return; }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
        float inv = 1.0f / (d23_1 + d23_2);
        s->v[1].a = d23_1 * inv; s->v[2].a = d23_2 * inv; s->count = 2; s->v[0] = s->v[2]; return;
    }
    float inv = 1.0f / (d123_1 + d123_2 + d123_3);
    s->v[0].a = d123_1 * inv; s->v[1].a = d123_2 * inv; s->v[2].a = d123_3 * inv; s->count = 3;
}

/* b2Distance(output, cache, input) with useRadii = false; returns output.distance */
static float gjk_distance(scache_t *cache, const dproxy_t *pA, xform xfA, const dproxy_t *pB, xform xfB)
{
    simplex_t sx;
    simplex_read_cache(&sx, cache, pA, xfA, pB, xfB);
    int saveA[3], saveB[3], saveCount = 0;
    int iter = 0;
    while (iter < 20) {
        saveCount = sx.count;
        for (int i = 0; i < saveCount; i++) { saveA[i] = sx.v[i].indexA; saveB[i] = sx.v[i].indexB; }
        if (sx.count == 2) simplex_solve2(&sx);
        else if (sx.count == 3) simplex_solve3(&sx);
        if (sx.count == 3) break;
        /* GetSearchDirection */
        v2 d;
        if (sx.count == 1) d = neg(sx.v[0].w);
        else {
            v2 e12 = sub(sx.v[1].w, sx.v[0].w);
            float sgn = crs(e12, neg(sx.v[0].w));
            d = sgn > 0.0f ? crs_sv(1.0f, e12) : crs_vs(e12, 1.0f);
        }
        if (dot(d, d) < B2_EPSILON * B2_EPSILON) break;
        sv_t *vx = &sx.v[sx.count];
        vx->indexA = proxy_support(pA, rmulT(xfA.q, neg(d)));
        vx->wA = xmul(xfA, pA->v[vx->indexA]);
        vx->indexB = proxy_support(pB, rmulT(xfB.q, d));
        vx->wB = xmul(xfB, pB->v[vx->indexB]);
        vx->w = sub(vx->wB, vx->wA);
        ++iter;
        int duplicate = 0;
        for (int i = 0; i < saveCount; i++) if (vx->indexA == saveA[i] && vx->indexB == saveB[i]) { duplicate = 1; break; }
        if (duplicate) break;
        ++sx.count;
    }
    /* GetWitnessPoints */
    v2 pointA, pointB;
    if (sx.count == 1) { pointA = sx.v[0].wA; pointB = sx.v[0].wB; }
    else if (sx.count == 2) {
        pointA = add(scl(sx.v[0].a, sx.v[0].wA), scl(sx.v[1].a, sx.v[1].wA));
        pointB = add(scl(sx.v[0].a, sx.v[0].wB), scl(sx.v[1].a, sx.v[1].wB));
    } else {
        pointA = add(add(scl(sx.v[0].a, sx.v[0].wA), scl(sx.v[1].a, sx.v[1].wA)), scl(sx.v[2].a, sx.v[2].wA));
        pointB = pointA;
    }
    simplex_write_cache(&sx, cache);
    return dist2(pointA, pointB);
}

/* ---------------------------------------------------------------- b2SeparationFunction */
typedef struct {
    const dproxy_t *pA, *pB;
    sweep_t sA, sB;
    int type; /* 0 points, 1 faceA, 2 faceB */
    v2 localPoint, axis;
} sepfn_t;

static v2 normalize_v(v2 v)
{ /* b2Vec2::Normalize (the vector is left unchanged when shorter than b2_epsilon) */
    float len = sqrtf(v.x * v.x + v.y * v.y);
    if (len < B2_EPSILON) return v;
    float inv = 1.0f / len;
    return V(v.x * inv, v.y * inv);
}

static void sepfn_init(sepfn_t *f, const scache_t *cache, const dproxy_t *pA, const sweep_t *sA, const dproxy_t *pB,
                       const sweep_t *sB, float t1)
{
    f->pA = pA; f->pB = pB; f->sA = *sA; f->sB = *sB;
    xform xfA = sweep_xf(&f->sA, t1), xfB = sweep_xf(&f->sB, t1);
    if (cache->count == 1) {
        f->type = 0;
        v2 pointA = xmul(xfA, pA->v[cache->indexA[0]]), pointB = xmul(xfB, pB->v[cache->indexB[0]]);
        f->axis = normalize_v(sub(pointB, pointA));
        f->localPoint = V(0.0f, 0.0f);
    } else if (cache->indexA[0] == cache->indexA[1]) {
        f->type = 2;
        v2 b1 = pB->v[cache->indexB[0]], b2 = pB->v[cache->indexB[1]];
        f->axis = normalize_v(crs_vs(sub(b2, b1), 1.0f));
        v2 normal = rmul(xfB.q, f->axis);
        f->localPoint = scl(0.5f, add(b1, b2));
        v2 pointB = xmul(xfB, f->localPoint), pointA = xmul(xfA, pA->v[cache->indexA[0]]);
        float s = dot(sub(pointA, pointB), normal);
        if (s < 0.0f) f->axis = neg(f->axis);
    } else {
        f->type = 1;
        v2 a1 = pA->v[cache->indexA[0]], a2 = pA->v[cache->indexA[1]];
        f->axis = normalize_v(crs_vs(sub(a2, a1), 1.0f));
        v2 normal = rmul(xfA.q, f->axis);
        f->localPoint = scl(0.5f, add(a1, a2));
        v2 pointA = xmul(xfA, f->localPoint), pointB = xmul(xfB, pB->v[cache->indexB[0]]);
        float s = dot(sub(pointB, pointA), normal);
        if (s < 0.0f) f->axis = neg(f->axis);
    }
}

/* FindMinSeparation (find = 1: picks the support indices) / Evaluate (find = 0: uses the given ones) */
static float sepfn_eval(const sepfn_t *f, int *indexA, int *indexB, float t, int find)
{
    xform xfA = sweep_xf(&f->sA, t), xfB = sweep_xf(&f->sB, t);
    if (f->type == 0) {
        if (find) {
            *indexA = proxy_support(f->pA, rmulT(xfA.q, f->axis));
            *indexB = proxy_support(f->pB, rmulT(xfB.q, neg(f->axis)));
        }
        v2 pointA = xmul(xfA, f->pA->v[*indexA]), pointB = xmul(xfB, f->pB->v[*indexB]);
        return dot(sub(pointB, pointA), f->axis);
    } else if (f->type == 1) {
        v2 normal = rmul(xfA.q, f->axis), pointA = xmul(xfA, f->localPoint);
        if (find) { *indexA = -1; *indexB = proxy_support(f->pB, rmulT(xfB.q, neg(normal))); }
        v2 pointB = xmul(xfB, f->pB->v[*indexB]);
        return dot(sub(pointB, pointA), normal);
    } else {
        v2 normal = rmul(xfB.q, f->axis), pointB = xmul(xfB, f->localPoint);
        if (find) { *indexB = -1; *indexA = proxy_support(f->pA, rmulT(xfA.q, neg(normal))); }
        v2 pointA = xmul(xfA, f->pA->v[*indexA]);
        return dot(sub(pointA, pointB), normal);
    }
}

/* ---------------------------------------------------------------- b2TimeOfImpact */
enum { TOI_UNKNOWN = 0, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };

static int time_of_impact(float *t_out, const dproxy_t *pA, const sweep_t *sweepA_, const dproxy_t *pB, const sweep_t *sweepB_,
                          float tMax)
{
    int state = TOI_UNKNOWN;
    *t_out = tMax;
    sweep_t sweepA = *sweepA_, sweepB = *sweepB_;
    sweep_normalize(&sweepA);
    sweep_normalize(&sweepB);
    const float totalRadius = POLYGON_RADIUS + POLYGON_RADIUS;
    const float target = fmaxf_(LINEAR_SLOP, totalRadius - 3.0f * LINEAR_SLOP);
    const float tolerance = 0.25f * LINEAR_SLOP;
    float t1 = 0.0f;
    int iter = 0;
    scache_t cache;
    cache.count = 0;
    for (;;) {
        xform xfA = sweep_xf(&sweepA, t1), xfB = sweep_xf(&sweepB, t1);
        float distance = gjk_distance(&cache, pA, xfA, pB, xfB);
        if (distance <= 0.0f) { state = TOI_OVERLAPPED; *t_out = 0.0f; break; }
        if (distance < target + tolerance) { state = TOI_TOUCHING; *t_out = t1; break; }
        sepfn_t fcn;
        sepfn_init(&fcn, &cache, pA, &sweepA, pB, &sweepB, t1);
        int done = 0;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            float s2 = sepfn_eval(&fcn, &indexA, &indexB, t2, 1);
            if (s2 > target + tolerance) { state = TOI_SEPARATED; *t_out = tMax; done = 1; break; }
            if (s2 > target - tolerance) { t1 = t2; break; }
            float s1 = sepfn_eval(&fcn, &indexA, &indexB, t1, 0);
            if (s1 < target - tolerance) { state = TOI_FAILED; *t_out = t1; done = 1; break; }
            if (s1 <= target + tolerance) { state = TOI_TOUCHING; *t_out = t1; done = 1; break; }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                ++rootIterCount;
                float s = sepfn_eval(&fcn, &indexA, &indexB, t, 0);
                if (fabsf(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                if (rootIterCount == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == 8 /* b2_maxPolygonVertices */) break;
        }
        ++iter;
        if (done) break;
        if (iter == 20) { state = TOI_FAILED; *t_out = t1; break; }
    }
    return state;
}

/* ---------------------------------------------------------------- b2World::SolveTOI */
typedef struct {
    int body, f;        /* f < ne: edge f; else static polygon f - ne */
    int sidx;           /* index into the static bodies' alpha0 table */
    int toiCount, toiValid, enabled;
    float toi;
} toicand_t;

static contact_t *pair_contact(b2l_world *W, int body, int f)
{
    return f < W->ne ? &W->ct[body * W->ne + f] : &W->ctp[body * W->np_cap + (f - W->ne)];
}
static float pair_friction(const b2l_world *W, int body, int f)
{
    return sqrtf((f < W->ne ? W->e[f].friction : W->sp[f - W->ne].friction) * W->b[body].friction);
}
static void pair_proxy(const b2l_world *W, int f, dproxy_t *p)
{
    if (f < W->ne) { p->count = 2; p->v[0] = W->e[f].v1; p->v[1] = W->e[f].v2; }
    else { const spoly_t *s = &W->sp[f - W->ne]; p->count = s->count; for (int i = 0; i < s->count; i++) p->v[i] = s->verts[i]; }
}
static int count_touching(const b2l_world *W)
{
    int n = 0;
    for (int b = 0; b < W->nb; b++) {
        for (int e = 0; e < W->ne; e++) n += W->ct[b * W->ne + e].touching;
        for (int p = 0; p < W->np; p++) n += W->ctp[b * W->np_cap + p].touching;
    }
    return n;
}

/* b2Contact::Update of one pair during SolveTOI (listener events included); the manifold-table capacity rule of
 * b2l_step applies */
static void toi_contact_update(b2l_world *W, int body, int f, int cap)
{
    contact_t *c = pair_contact(W, body, f);
    int was = c->touching;
    int ev = f < W->ne ? contact_update(c, &W->e[f], &W->b[body]) : contact_update_poly(c, &W->sp[f - W->ne], &W->b[body]);
    if (c->touching && !was && count_touching(W) > cap) { c->touching = 0; c->m.pointCount = 0; W->overflowed++; ev = 0; }
    if (ev != 0 && W->event) W->event(W->ctx, body, ev > 0);
}

/* fat AABB of a body's polygon over its sweep (b2Body::SynchronizeFixtures: the union of the AABBs at the
 * sweep's start transform and at the current one, + b2_aabbExtension) */
static void swept_aabb(const body_t *b, float *lox, float *loy, float *hix, float *hiy)
{
    sweep_t s;
    s.localCenter = b->localCenter; s.c0 = b->c0; s.c = b->c; s.a0 = b->a0; s.a = b->a; s.alpha0 = b->alpha0;
    xform xf0 = sweep_xf(&s, 0.0f);
    *lox = 3.402823466e+38f; *loy = *lox; *hix = -*lox; *hiy = -*lox;
    for (int i = 0; i < b->count; i++) {
        v2 p = xmul(xf0, b->verts[i]), q = xmul(b->xf, b->verts[i]);
        *lox = fminf_(*lox, fminf_(p.x, q.x)); *loy = fminf_(*loy, fminf_(p.y, q.y));
        *hix = fmaxf_(*hix, fmaxf_(p.x, q.x)); *hiy = fmaxf_(*hiy, fmaxf_(p.y, q.y));
    }
}

static int toi_find(const toicand_t *cand, int n, int body, int f)
{
    for (int i = 0; i < n; i++) if (cand[i].body == body && cand[i].f == f) return i;
    return -1;
}

/* b2ContactManager::FindNewContacts for one body: pairs whose fat AABBs overlap join the contact list */
static void toi_add_candidates(b2l_world *W, toicand_t *cand, int *ncand, float *alphaS, int *nstat, int *statId, int body)
{
    const body_t *b = &W->b[body];
    float lox, loy, hix, hiy;
    swept_aabb(b, &lox, &loy, &hix, &hiy);
    const float ext = POLYGON_RADIUS + AABB_EXTENSION;
    for (int f = W->ne + W->np - 1; f >= 0; f--) {
        float elox, ehix, eloy, ehiy;
        if (f < W->ne) {
            const edge_t *e = &W->e[f];
            elox = fminf_(e->v1.x, e->v2.x) - ext; ehix = fmaxf_(e->v1.x, e->v2.x) + ext;
            eloy = fminf_(e->v1.y, e->v2.y) - ext; ehiy = fmaxf_(e->v1.y, e->v2.y) + ext;
        } else {
            const spoly_t *sp = &W->sp[f - W->ne];
            elox = sp->verts[2].x - ext; ehix = sp->verts[0].x + ext; eloy = sp->verts[0].y - ext; ehiy = sp->verts[2].y + ext;
        }
        if (lox - ext > ehix || elox > hix + ext || loy - ext > ehiy || eloy > hiy + ext) continue;
        if (toi_find(cand, *ncand, body, f) >= 0) continue;
        if (*ncand >= B2L_MAX_TOI_CAND) { W->overflowed++; continue; }
        toicand_t *c = &cand[(*ncand)++];
        c->body = body; c->f = f; c->toiCount = 0; c->toiValid = 0; c->enabled = 1; c->toi = 1.0f;
        const int sid = W->one_static_body ? 0 : f;
        int si = -1;
        for (int k = 0; k < *nstat; k++) if (statId[k] == sid) { si = k; break; }
        if (si < 0) { si = (*nstat)++; statId[si] = sid; alphaS[si] = 0.0f; }
        c->sidx = si;
    }
}

/* b2Body::Advance */
static void body_advance(body_t *b, float alpha)
{
    sweep_t s;
    s.localCenter = b->localCenter; s.c0 = b->c0; s.c = b->c; s.a0 = b->a0; s.a = b->a; s.alpha0 = b->alpha0;
    sweep_advance(&s, alpha);
    b->c0 = s.c0; b->a0 = s.a0; b->alpha0 = s.alpha0;
    b->c = b->c0; b->a = b->a0;
    body_sync_xf(b);
}

/* b2Island::SolveTOI for the island {static bodies, body B}: the contacts all have B as their only movable body */
static void island_solve_toi(b2l_world *W, int body, const int *ic, int nic, const toicand_t *cand, float h, int velIters)
{
    body_t *B = &W->b[body];
    const float mB = B->invMass, iB = B->invI;
    v2 cB = B->c, vB = B->v;
    float aB = B->a, wB = B->w;
    /* b2ContactSolver: position constraints hold the manifolds as they are now */
    const manifold_t *mf[B2L_MAX_CONTACTS];
    float fric[B2L_MAX_CONTACTS];
    for (int k = 0; k < nic; k++) { mf[k] = &pair_contact(W, body, cand[ic[k]].f)->m; fric[k] = pair_friction(W, body, cand[ic[k]].f); }
    /* SolveTOIPositionConstraints, <= 20 iterations */
    for (int it = 0; it < 20; it++) {
        float minSep = 0.0f;
        for (int k = 0; k < nic; k++) {
            const manifold_t *m = mf[k];
            for (int p = 0; p < m->pointCount; p++) {
                xform xfB;
                xfB.q = rot_of(aB);
                xfB.p = sub(cB, rmul(xfB.q, B->localCenter));
                v2 normal, point;
                float separation;
                if (m->type == 0) {
                    normal = m->localNormal;
                    v2 plane = m->localPoint;
                    v2 clip = xmul(xfB, m->pts[p].localPoint);
                    separation = dot(sub(clip, plane), normal) - POLYGON_RADIUS - POLYGON_RADIUS;
                    point = clip;
                } else {
                    normal = rmul(xfB.q, m->localNormal);
                    v2 plane = xmul(xfB, m->localPoint);
                    v2 clip = m->pts[p].localPoint;
                    separation = dot(sub(clip, plane), normal) - POLYGON_RADIUS - POLYGON_RADIUS;
                    point = clip;
                    normal = neg(normal);
                }
                v2 rB = sub(point, cB);
                minSep = fminf_(minSep, separation);
                float C = clampf(TOI_BAUMGARTE * (separation + LINEAR_SLOP), -MAX_LINEAR_CORRECTION, 0.0f);
                float rnB = crs(rB, normal);
                float K = mB + iB * rnB * rnB;
                float impulse = K > 0.0f ? -C / K : 0.0f;
                v2 P = scl(impulse, normal);
                cB = add(cB, scl(mB, P));
                aB += iB * crs(rB, P);
            }
        }
        if (minSep >= -1.5f * LINEAR_SLOP) break;
    }
    /* leap of faith to the new safe state */
    B->c0 = cB; B->a0 = aB;
    /* InitializeVelocityConstraints (no warm starting: impulses start from zero) */
    vc_t vc[B2L_MAX_CONTACTS];
    for (int k = 0; k < nic; k++) {
        vc_t *q = &vc[k];
        q->pointCount = mf[k]->pointCount; q->friction = fric[k];
        xform xfB;
        xfB.q = rot_of(aB);
        xfB.p = sub(cB, rmul(xfB.q, B->localCenter));
        v2 pts[2];
        world_manifold(mf[k], XF_ID, xfB, POLYGON_RADIUS, POLYGON_RADIUS, &q->normal, pts);
        for (int p = 0; p < q->pointCount; p++) {
            vcp_t *cp = &q->p[p];
            cp->normalImpulse = 0.0f; cp->tangentImpulse = 0.0f;
            cp->rA = pts[p];
            cp->rB = sub(pts[p], cB);
            float rnB = crs(cp->rB, q->normal);
            float kN = mB + iB * rnB * rnB;
            cp->normalMass = kN > 0.0f ? 1.0f / kN : 0.0f;
            v2 tangent = crs_vs(q->normal, 1.0f);
            float rtB = crs(cp->rB, tangent);
            float kT = mB + iB * rtB * rtB;
            cp->tangentMass = kT > 0.0f ? 1.0f / kT : 0.0f;
            cp->velocityBias = 0.0f;
        }
        if (q->pointCount == 2) {
            float rn1B = crs(q->p[0].rB, q->normal), rn2B = crs(q->p[1].rB, q->normal);
            float k11 = mB + iB * rn1B * rn1B, k22 = mB + iB * rn2B * rn2B, k12 = mB + iB * rn1B * rn2B;
            if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
                q->K[0][0] = k11; q->K[0][1] = k12; q->K[1][0] = k12; q->K[1][1] = k22;
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = 1.0f / det;
                q->nM[0][0] = det * k22; q->nM[1][0] = -det * k12; q->nM[0][1] = -det * k12; q->nM[1][1] = det * k11;
            } else q->pointCount = 1;
        }
    }
    /* SolveVelocityConstraints x velocityIterations */
    for (int it = 0; it < velIters; it++) {
        for (int k = 0; k < nic; k++) {
            vc_t *q = &vc[k];
            v2 normal = q->normal, tangent = crs_vs(normal, 1.0f);
            for (int p = 0; p < q->pointCount; p++) {
                vcp_t *cp = &q->p[p];
                v2 dv = add(vB, crs_sv(wB, cp->rB));
                float vt = dot(dv, tangent) - 0.0f;
                float lambda = cp->tangentMass * (-vt);
                float maxF = q->friction * cp->normalImpulse;
                float newImp = clampf(cp->tangentImpulse + lambda, -maxF, maxF);
                lambda = newImp - cp->tangentImpulse;
                cp->tangentImpulse = newImp;
                v2 P = scl(lambda, tangent);
                vB = add(vB, scl(mB, P)); wB += iB * crs(cp->rB, P);
            }
            if (q->pointCount == 1) {
                vcp_t *cp = &q->p[0];
                v2 dv = add(vB, crs_sv(wB, cp->rB));
                float vn = dot(dv, normal);
                float lambda = -cp->normalMass * (vn - cp->velocityBias);
                float newImp = fmaxf_(cp->normalImpulse + lambda, 0.0f);
                lambda = newImp - cp->normalImpulse;
                cp->normalImpulse = newImp;
                v2 P = scl(lambda, normal);
                vB = add(vB, scl(mB, P)); wB += iB * crs(cp->rB, P);
            } else {
                vcp_t *c1 = &q->p[0], *c2 = &q->p[1];
                v2 a = V(c1->normalImpulse, c2->normalImpulse);
                v2 dv1 = add(vB, crs_sv(wB, c1->rB)), dv2 = add(vB, crs_sv(wB, c2->rB));
                float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
                v2 b = V(vn1 - c1->velocityBias, vn2 - c2->velocityBias);
                b = sub(b, V(q->K[0][0] * a.x + q->K[1][0] * a.y, q->K[0][1] * a.x + q->K[1][1] * a.y));
                v2 x;
                int solved = 0;
                x = V(-(q->nM[0][0] * b.x + q->nM[1][0] * b.y), -(q->nM[0][1] * b.x + q->nM[1][1] * b.y));
                if (x.x >= 0.0f && x.y >= 0.0f) solved = 1;
                if (!solved) {
                    x.x = -c1->normalMass * b.x; x.y = 0.0f;
                    vn2 = q->K[0][1] * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) solved = 1;
                }
                if (!solved) {
                    x.x = 0.0f; x.y = -c2->normalMass * b.y;
                    vn1 = q->K[1][0] * x.y + b.x;
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
        }
    }
    /* the TOI impulses are NOT stored for warm starting; integrate positions over the rest of the step */
    {
        v2 tr = scl(h, vB);
        if (dot(tr, tr) > MAX_TRANSLATION * MAX_TRANSLATION) { float ratio = MAX_TRANSLATION / sqrtf(dot(tr, tr)); vB = scl(ratio, vB); }
        float rotn = h * wB;
        if (rotn * rotn > MAX_ROTATION * MAX_ROTATION) { float ratio = MAX_ROTATION / fabsf(rotn); wB *= ratio; }
        cB = add(cB, scl(h, vB));
        aB = aB + h * wB;
    }
    B->c = cB; B->a = aB; B->v = vB; B->w = wB;
    body_sync_xf(B);
}

static void b2l_solve_toi(b2l_world *W, float dt, int velIters)
{
    const int NB = W->nb;
    const int cap = (W->max_contacts > 0 && W->max_contacts < B2L_MAX_CONTACTS) ? W->max_contacts : B2L_MAX_CONTACTS;
    toicand_t cand[B2L_MAX_TOI_CAND];
    float alphaS[B2L_MAX_TOI_CAND];
    int statId[B2L_MAX_TOI_CAND];
    int ncand = 0, nstat = 0;
    W->stat_toi_events = 0; W->stat_toi_calls = 0;
    for (int i = 0; i < NB; i++) W->b[i].alpha0 = 0.0f;
    for (int oi = 0; oi < NB; oi++) toi_add_candidates(W, cand, &ncand, alphaS, &nstat, statId, W->body_order[oi]);
    for (;;) {
        int minIdx = -1;
        float minAlpha = 1.0f;
        for (int ci = 0; ci < ncand; ci++) {
            toicand_t *c = &cand[ci];
            if (!c->enabled) continue;
            if (c->toiCount > MAX_SUB_STEPS) continue;
            float alpha = 1.0f;
            if (c->toiValid) alpha = c->toi;
            else {
                body_t *B = &W->b[c->body];
                if (!B->awake) continue;
                float alpha0 = alphaS[c->sidx];
                if (alphaS[c->sidx] < B->alpha0) { alpha0 = B->alpha0; alphaS[c->sidx] = alpha0; }
                else if (B->alpha0 < alphaS[c->sidx]) {
                    alpha0 = alphaS[c->sidx];
                    sweep_t s;
                    s.localCenter = B->localCenter; s.c0 = B->c0; s.c = B->c; s.a0 = B->a0; s.a = B->a; s.alpha0 = B->alpha0;
                    sweep_advance(&s, alpha0);
                    B->c0 = s.c0; B->a0 = s.a0; B->alpha0 = s.alpha0;
                }
                dproxy_t pA, pB;
                pair_proxy(W, c->f, &pA);
                pB.count = B->count;
                for (int i = 0; i < B->count; i++) pB.v[i] = B->verts[i];
                sweep_t sA, sB;
                sA.localCenter = V(0.0f, 0.0f); sA.c0 = V(0.0f, 0.0f); sA.c = V(0.0f, 0.0f); sA.a0 = 0.0f; sA.a = 0.0f; sA.alpha0 = alpha0;
                sB.localCenter = B->localCenter; sB.c0 = B->c0; sB.c = B->c; sB.a0 = B->a0; sB.a = B->a; sB.alpha0 = B->alpha0;
                float beta;
                int state = time_of_impact(&beta, &pA, &sA, &pB, &sB, 1.0f);
                W->stat_toi_calls++;
                if (state == TOI_TOUCHING) alpha = fminf_(alpha0 + (1.0f - alpha0) * beta, 1.0f);
                else alpha = 1.0f;
                c->toi = alpha; c->toiValid = 1;
            }
            if (alpha < minAlpha) { minIdx = ci; minAlpha = alpha; }
        }
        if (minIdx < 0 || 1.0f - 10.0f * B2_EPSILON < minAlpha) break;
        toicand_t *mc = &cand[minIdx];
        body_t *B = &W->b[mc->body];
        /* advance the bodies to the TOI */
        const float backupS = alphaS[mc->sidx];
        const body_t backupB = *B;
        alphaS[mc->sidx] = minAlpha;
        body_advance(B, minAlpha);
        toi_contact_update(W, mc->body, mc->f, cap);
        mc->toiValid = 0;
        ++mc->toiCount;
        if (!pair_contact(W, mc->body, mc->f)->touching) {
            mc->enabled = 0;
            alphaS[mc->sidx] = backupS;
            B->c0 = backupB.c0; B->c = backupB.c; B->a0 = backupB.a0; B->a = backupB.a; B->alpha0 = backupB.alpha0;
            body_sync_xf(B);
            continue;
        }
        /* build the island: the TOI contact, then the body's other touching contacts against static bodies */
        int ic[B2L_MAX_CONTACTS], nic = 0;
        int inIsland[B2L_MAX_TOI_CAND]; /* static bodies (by sidx) already in the island */
        for (int k = 0; k < nstat; k++) inIsland[k] = 0;
        ic[nic++] = minIdx;
        inIsland[mc->sidx] = 1;
        for (int ci = 0; ci < ncand; ci++) {
            toicand_t *oc = &cand[ci];
            if (ci == minIdx || oc->body != mc->body) continue;
            if (nic == B2L_MAX_CONTACTS) break;
            const float backup = alphaS[oc->sidx];
            if (!inIsland[oc->sidx]) alphaS[oc->sidx] = minAlpha;
            toi_contact_update(W, oc->body, oc->f, cap);
            oc->enabled = 1; /* b2Contact::Update re-enables the contact */
            if (!pair_contact(W, oc->body, oc->f)->touching) { alphaS[oc->sidx] = backup; continue; }
            ic[nic++] = ci;
            inIsland[oc->sidx] = 1;
        }
        W->stat_toi_events++;
        island_solve_toi(W, mc->body, ic, nic, cand, (1.0f - minAlpha) * dt, velIters);
        /* invalidate all contact TOIs on the displaced body; its moved proxy may create new contacts */
        for (int ci = 0; ci < ncand; ci++) if (cand[ci].body == mc->body) cand[ci].toiValid = 0;
        toi_add_candidates(W, cand, &ncand, alphaS, &nstat, statId, mc->body);
    }
}

#endif
