// b2lite_toi.cuh -- continuous collision of Box2D 2.3 on the device (included by b2lite.cuh, inside namespace b2l).
//
// b2World::SolveTOI for the one case gym's Box2D tasks exercise: a non-bullet dynamic polygon against fixtures of
// static bodies (terrain edges; BipedalWalkerHardcore's boxes).  Per step, after the discrete solve:
//   * the contact list = every (body, static fixture) pair whose fat AABBs overlap, the body's AABB taken over its
//     sweep (b2Body::SynchronizeFixtures); visited like every other per-pair loop here (island order of the bodies x
//     descending fixture index);
//   * b2TimeOfImpact per pair: conservative advancement on a separating axis (b2SeparationFunction) seeded by GJK
//     (b2Distance with a simplex cache), target separation = linearSlop, tolerance linearSlop / 4, <= 20 outer
//     iterations, <= 8 push-backs, <= 50 root-finder steps (bisection / secant);
//   * the earliest event: both sweeps advanced to it, the manifold re-evaluated (Begin/EndContact events fire here,
//     one step earlier than the discrete Collide would see them), a mini-island {the body, its touching contacts
//     against static bodies}: b2ContactSolver::SolveTOIPositionConstraints (Baumgarte 0.75, <= 20 iterations, until
//     -1.5 linearSlop), the step's 180 velocity iterations without warm starting, integration over the rest of the
//     step; impulses are not stored; joints are not part of the TOI island (as in Box2D);
//   * repeat until no pair has an event left (each pair at most 9 times: b2_maxSubSteps = 8).
// The static bodies' sweeps only carry alpha0 (Box2D advances it with the dynamic body's, and a later pair of the
// same static body starts its interval there): ONE static body owns every edge of LunarLander's moon, ONE static
// body per fixture in BipedalWalker (Scene::kOneStaticBody).
// float32, one rounding per operation, the same operation order as the CPU checker in the test tree.
#pragma once

// workload counters of the host build (tests/hostsim with -DB2L_TOI_STATS): {pairs examined, skipped by the AABB
// shortcut, skipped by the resting-contact shortcut, b2TimeOfImpact evaluations, TOI sub-steps, velocity iterations run}
#if defined(B2L_TOI_STATS) && !defined(__CUDA_ARCH__)
extern long long b2l_toi_stats[6];
#define B2L_STAT(i, n) (b2l_toi_stats[i] += (n))
#else
#define B2L_STAT(i, n) ((void)0)
#endif

constexpr float kEpsilon = 1.1920929e-07f;
constexpr float kPi = 3.14159265359f;
constexpr int kMaxSubSteps = 8;
constexpr float kToiBaumgarte = 0.75f;
constexpr int kMaxToiCand = 48;   // contact-list capacity per step (5 bodies x the ~8 fixtures a fallen walker can span)

// ---- b2Sweep -----------------------------------------------------------------------------------------
struct Sweep { v2 localCenter, c0, c; float a0, a, alpha0; };

LD xform sweep_xf(const Sweep &s, float beta) {  // b2Sweep::GetTransform
    xform xf;
    xf.p = add(scl(1.0f - beta, s.c0), scl(beta, s.c));
    const float angle = (1.0f - beta) * s.a0 + beta * s.a;
    xf.q = rot_of(angle);
    xf.p = sub(xf.p, rmul(xf.q, s.localCenter));
    return xf;
}
LD void sweep_advance(Sweep &s, float alpha) {  // b2Sweep::Advance
    const float beta = (alpha - s.alpha0) / (1.0f - s.alpha0);
    s.c0 = add(s.c0, scl(beta, sub(s.c, s.c0)));
    s.a0 = s.a0 + beta * (s.a - s.a0);
    s.alpha0 = alpha;
}
LD void sweep_normalize(Sweep &s) {  // b2Sweep::Normalize
    const float twoPi = 2.0f * kPi;
    const float d = twoPi * floorf(s.a0 / twoPi);
    s.a0 -= d;
    s.a -= d;
}

// ---- b2Distance (GJK) --------------------------------------------------------------------------------
struct DProxy { int count; v2 v[MAXV]; };
struct SV { v2 wA, wB, w; float a; int indexA, indexB; };
struct Simplex { SV v[3]; int count; };
struct SCache { float metric; int count; int indexA[3], indexB[3]; };

LD int proxy_support(const DProxy &p, v2 d) {
    int best = 0;
    float bestValue = dot(p.v[0], d);
    for (int i = 1; i < p.count; i++) { const float value = dot(p.v[i], d); if (value > bestValue) { best = i; bestValue = value; } }
    return best;
}
LD float dist2(v2 a, v2 b) { const v2 c = sub(a, b); return sqrtf(c.x * c.x + c.y * c.y); }

LD float simplex_metric(const Simplex &s) {
    if (s.count == 2) return dist2(s.v[0].w, s.v[1].w);
    if (s.count == 3) return crs(sub(s.v[1].w, s.v[0].w), sub(s.v[2].w, s.v[0].w));
    return 0.0f;
}
LD void simplex_solve2(Simplex &s) {
    const v2 w1 = s.v[0].w, w2 = s.v[1].w, e12 = sub(w2, w1);
    const float d12_2 = -dot(w1, e12);
    if (d12_2 <= 0.0f) { s.v[0].a = 1.0f; s.count = 1; return; }
    const float d12_1 = dot(w2, e12);
    if (d12_1 <= 0.0f) { s.v[1].a = 1.0f; s.count = 1; s.v[0] = s.v[1]; return; }
    const float inv_d12 = 1.0f / (d12_1 + d12_2);
    s.v[0].a = d12_1 * inv_d12; s.v[1].a = d12_2 * inv_d12; s.count = 2;
}
LD void simplex_solve3(Simplex &s) {
    const v2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;
    const v2 e12 = sub(w2, w1);
    const float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12), d12_1 = w2e12, d12_2 = -w1e12;
    const v2 e13 = sub(w3, w1);
    const float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13), d13_1 = w3e13, d13_2 = -w1e13;
    const v2 e23 = sub(w3, w2);
    const float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23), d23_1 = w3e23, d23_2 = -w2e23;
    const float n123 = crs(e12, e13);
    const float d123_1 = n123 * crs(w2, w3), d123_2 = n123 * crs(w3, w1), d123_3 = n123 * crs(w1, w2);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) { s.v[0].a = 1.0f; s.count = 1; return; }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
        const float inv = 1.0f / (d12_1 + d12_2);
        s.v[0].a = d12_1 * inv; s.v[1].a = d12_2 * inv; s.count = 2; return;
    }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
        const float inv = 1.0f / (d13_1 + d13_2);
        s.v[0].a = d13_1 * inv; s.v[2].a = d13_2 * inv; s.count = 2; s.v[1] = s.v[2]; return;
    }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) { s.v[1].a = 1.0f; s.count = 1; s.v[0] = s.v[1]; return; }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) { s.v[2].a = 1.0f; s.count = 1; s.v[0] = s.v[2]; return; }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
        const float inv = 1.0f / (d23_1 + d23_2);
        s.v[1].a = d23_1 * inv; s.v[2].a = d23_2 * inv; s.count = 2; s.v[0] = s.v[2]; return;
    }
    const float inv = 1.0f / (d123_1 + d123_2 + d123_3);
    s.v[0].a = d123_1 * inv; s.v[1].a = d123_2 * inv; s.v[2].a = d123_3 * inv; s.count = 3;
}

// b2Distance(output, cache, input) with useRadii = false; returns output.distance
__device__ __noinline__ float gjk_distance(SCache &cache, const DProxy &pA, const xform &xfA, const DProxy &pB, const xform &xfB) {
    Simplex sx;
    // ReadCache
    sx.count = cache.count;
    for (int i = 0; i < sx.count; i++) {
        SV &v = sx.v[i];
        v.indexA = cache.indexA[i]; v.indexB = cache.indexB[i];
        v.wA = xmul(xfA, pA.v[v.indexA]); v.wB = xmul(xfB, pB.v[v.indexB]);
        v.w = sub(v.wB, v.wA); v.a = 0.0f;
    }
    if (sx.count > 1) {
        const float metric1 = cache.metric, metric2 = simplex_metric(sx);
        if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) sx.count = 0;
    }
    if (sx.count == 0) {
        SV &v = sx.v[0];
        v.indexA = 0; v.indexB = 0;
        v.wA = xmul(xfA, pA.v[0]); v.wB = xmul(xfB, pB.v[0]);
        v.w = sub(v.wB, v.wA); v.a = 1.0f;
        sx.count = 1;
    }
    int saveA[3], saveB[3], saveCount = 0;
    int iter = 0;
    while (iter < 20) {
        saveCount = sx.count;
        for (int i = 0; i < saveCount; i++) { saveA[i] = sx.v[i].indexA; saveB[i] = sx.v[i].indexB; }
        if (sx.count == 2) simplex_solve2(sx);
        else if (sx.count == 3) simplex_solve3(sx);
        if (sx.count == 3) break;
        v2 d;  // GetSearchDirection
        if (sx.count == 1) d = neg(sx.v[0].w);
        else {
            const v2 e12 = sub(sx.v[1].w, sx.v[0].w);
            const float sgn = crs(e12, neg(sx.v[0].w));
            d = sgn > 0.0f ? crs_sv(1.0f, e12) : crs_vs(e12, 1.0f);
        }
        if (dot(d, d) < kEpsilon * kEpsilon) break;
        SV &vx = sx.v[sx.count];
        vx.indexA = proxy_support(pA, rmulT(xfA.q, neg(d)));
        vx.wA = xmul(xfA, pA.v[vx.indexA]);
        vx.indexB = proxy_support(pB, rmulT(xfB.q, d));
        vx.wB = xmul(xfB, pB.v[vx.indexB]);
        vx.w = sub(vx.wB, vx.wA);
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; i++) if (vx.indexA == saveA[i] && vx.indexB == saveB[i]) { duplicate = true; break; }
        if (duplicate) break;
        ++sx.count;
    }
    v2 pointA, pointB;  // GetWitnessPoints
    if (sx.count == 1) { pointA = sx.v[0].wA; pointB = sx.v[0].wB; }
    else if (sx.count == 2) {
        pointA = add(scl(sx.v[0].a, sx.v[0].wA), scl(sx.v[1].a, sx.v[1].wA));
        pointB = add(scl(sx.v[0].a, sx.v[0].wB), scl(sx.v[1].a, sx.v[1].wB));
    } else {
        pointA = add(add(scl(sx.v[0].a, sx.v[0].wA), scl(sx.v[1].a, sx.v[1].wA)), scl(sx.v[2].a, sx.v[2].wA));
        pointB = pointA;
    }
    cache.metric = simplex_metric(sx);  // WriteCache
    cache.count = sx.count;
    for (int i = 0; i < sx.count; i++) { cache.indexA[i] = sx.v[i].indexA; cache.indexB[i] = sx.v[i].indexB; }
    return dist2(pointA, pointB);
}

// ---- b2SeparationFunction ------------------------------------------------------------------------------
struct SepFn {
    const DProxy *pA, *pB;
    Sweep sA, sB;
    int type;  // 0 points, 1 faceA, 2 faceB
    v2 localPoint, axis;
};

LD v2 normalize_v(v2 v) {  // b2Vec2::Normalize (left unchanged when shorter than b2_epsilon)
    const float len = sqrtf(v.x * v.x + v.y * v.y);
    if (len < kEpsilon) return v;
    const float inv = 1.0f / len;
    return V(v.x * inv, v.y * inv);
}

LD void sepfn_init(SepFn &f, const SCache &cache, const DProxy &pA, const Sweep &sA, const DProxy &pB, const Sweep &sB, float t1) {
    f.pA = &pA; f.pB = &pB; f.sA = sA; f.sB = sB;
    const xform xfA = sweep_xf(f.sA, t1), xfB = sweep_xf(f.sB, t1);
    if (cache.count == 1) {
        f.type = 0;
        const v2 pointA = xmul(xfA, pA.v[cache.indexA[0]]), pointB = xmul(xfB, pB.v[cache.indexB[0]]);
        f.axis = normalize_v(sub(pointB, pointA));
        f.localPoint = V(0.0f, 0.0f);
    } else if (cache.indexA[0] == cache.indexA[1]) {
        f.type = 2;
        const v2 b1 = pB.v[cache.indexB[0]], b2 = pB.v[cache.indexB[1]];
        f.axis = normalize_v(crs_vs(sub(b2, b1), 1.0f));
        const v2 normal = rmul(xfB.q, f.axis);
        f.localPoint = scl(0.5f, add(b1, b2));
        const v2 pointB = xmul(xfB, f.localPoint), pointA = xmul(xfA, pA.v[cache.indexA[0]]);
        const float s = dot(sub(pointA, pointB), normal);
        if (s < 0.0f) f.axis = neg(f.axis);
    } else {
        f.type = 1;
        const v2 a1 = pA.v[cache.indexA[0]], a2 = pA.v[cache.indexA[1]];
        f.axis = normalize_v(crs_vs(sub(a2, a1), 1.0f));
        const v2 normal = rmul(xfA.q, f.axis);
        f.localPoint = scl(0.5f, add(a1, a2));
        const v2 pointA = xmul(xfA, f.localPoint), pointB = xmul(xfB, pB.v[cache.indexB[0]]);
        const float s = dot(sub(pointB, pointA), normal);
        if (s < 0.0f) f.axis = neg(f.axis);
    }
}

// FindMinSeparation (find: picks the support indices) / Evaluate (uses the given ones)
__device__ __noinline__ float sepfn_eval(const SepFn &f, int &indexA, int &indexB, float t, bool find) {
    const xform xfA = sweep_xf(f.sA, t), xfB = sweep_xf(f.sB, t);
    if (f.type == 0) {
        if (find) {
            indexA = proxy_support(*f.pA, rmulT(xfA.q, f.axis));
            indexB = proxy_support(*f.pB, rmulT(xfB.q, neg(f.axis)));
        }
        const v2 pointA = xmul(xfA, f.pA->v[indexA]), pointB = xmul(xfB, f.pB->v[indexB]);
        return dot(sub(pointB, pointA), f.axis);
    } else if (f.type == 1) {
        const v2 normal = rmul(xfA.q, f.axis), pointA = xmul(xfA, f.localPoint);
        if (find) { indexA = -1; indexB = proxy_support(*f.pB, rmulT(xfB.q, neg(normal))); }
        const v2 pointB = xmul(xfB, f.pB->v[indexB]);
        return dot(sub(pointB, pointA), normal);
    } else {
        const v2 normal = rmul(xfB.q, f.axis), pointB = xmul(xfB, f.localPoint);
        if (find) { indexB = -1; indexA = proxy_support(*f.pA, rmulT(xfA.q, neg(normal))); }
        const v2 pointA = xmul(xfA, f.pA->v[indexA]);
        return dot(sub(pointA, pointB), normal);
    }
}

// ---- b2TimeOfImpact ------------------------------------------------------------------------------------
enum { TOI_UNKNOWN = 0, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };

__device__ __noinline__ int time_of_impact(float &t_out, const DProxy &pA, const Sweep &sweepA_, const DProxy &pB,
                                           const Sweep &sweepB_, float tMax) {
    int state = TOI_UNKNOWN;
    t_out = tMax;
    Sweep sweepA = sweepA_, sweepB = sweepB_;
    sweep_normalize(sweepA);
    sweep_normalize(sweepB);
    const float totalRadius = kPolygonRadius + kPolygonRadius;
    const float target = fmax_(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    const float tolerance = 0.25f * kLinearSlop;
    float t1 = 0.0f;
    int iter = 0;
    SCache cache;
    cache.count = 0;
    for (;;) {
        const xform xfA = sweep_xf(sweepA, t1), xfB = sweep_xf(sweepB, t1);
        const float distance = gjk_distance(cache, pA, xfA, pB, xfB);
        if (distance <= 0.0f) { state = TOI_OVERLAPPED; t_out = 0.0f; break; }
        if (distance < target + tolerance) { state = TOI_TOUCHING; t_out = t1; break; }
        SepFn fcn;
        sepfn_init(fcn, cache, pA, sweepA, pB, sweepB, t1);
        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            float s2 = sepfn_eval(fcn, indexA, indexB, t2, true);
            if (s2 > target + tolerance) { state = TOI_SEPARATED; t_out = tMax; done = true; break; }
            if (s2 > target - tolerance) { t1 = t2; break; }
            float s1 = sepfn_eval(fcn, indexA, indexB, t1, false);
            if (s1 < target - tolerance) { state = TOI_FAILED; t_out = t1; done = true; break; }
            if (s1 <= target + tolerance) { state = TOI_TOUCHING; t_out = t1; done = true; break; }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5f * (a1 + a2);
                ++rootIterCount;
                const float s = sepfn_eval(fcn, indexA, indexB, t, false);
                if (fabsf(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                if (rootIterCount == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == 8 /* b2_maxPolygonVertices */) break;
        }
        ++iter;
        if (done) break;
        if (iter == 20) { state = TOI_FAILED; t_out = t1; break; }
    }
    return state;
}

// ---- b2World::SolveTOI ---------------------------------------------------------------------------------
struct ToiCand { int body, f, sidx, toiCount; bool toiValid, enabled; float toi; };

LD Sweep body_sweep(const Body &B, const ShapeConst &sh) {
    Sweep s;
    s.localCenter = sh.localCenter; s.c0 = B.c0; s.c = B.c; s.a0 = B.a0; s.a = B.a; s.alpha0 = B.alpha0;
    return s;
}

// the distance proxy of static fixture f: an edge's two vertices, or a box's four corners in b2PolygonShape::Set order
template <typename Scene>
LD void fixture_proxy(const typename Scene::World &W, int f, DProxy &pA) {
    constexpr int NE = Scene::NE;
    if (Scene::NP > 0 && f >= NE) {
        if constexpr (Scene::NP > 0) {
            float x0, ylo, x1, yhi, fr;
            Scene::poly(W, f - NE, x0, ylo, x1, yhi, fr);
            pA.count = 4;
            pA.v[0] = V(x1, ylo); pA.v[1] = V(x1, yhi); pA.v[2] = V(x0, yhi); pA.v[3] = V(x0, ylo);
        }
    } else {
        float fr;
        pA.count = 2;
        Scene::edge(W, f, pA.v[0], pA.v[1], fr);
    }
}

// true when b2TimeOfImpact of the polygon `sh` sweeping along sB (current transform xf1 = the sweep's end) against
// the static fixture pA provably cannot report e_touching, i.e. the pair's alpha is 1 without running it.
__device__ __noinline__ bool toi_cannot_touch(const ShapeConst &sh, const xform &xf1, const Sweep &sB, const DProxy &pA) {
    // Shortcut (results unchanged): b2TimeOfImpact can only report e_touching if some transform of the
    // sweep brings the core shapes within target + tolerance = 1.25 linearSlop.  Every vertex moves from
    // its start to its end position within R * (1 - cos(d/2)) <= R * d^2 / 8 of the straight segment
    // between the two (R: its distance from the centre of mass, d: the rotation over the sweep), so the
    // polygon never leaves the box around its start and end poses widened by that amount.  If that box,
    // widened once more by 1.25 linearSlop and a safety margin far above float32 rounding, misses the
    // fixture's box, every outcome is "separated" (or "failed"), i.e. alpha = 1.
    bool far_apart = false;
    {
        const xform xf0 = sweep_xf(sB, 0.0f);
        float lox = kFltMax, loy = kFltMax, hix = -kFltMax, hiy = -kFltMax, r2 = 0.0f;
        v2 p0[MAXV];
        for (int i = 0; i < sh.count; i++) {
            const v2 p = xmul(xf0, sh.verts[i]), q = xmul(xf1, sh.verts[i]);
            p0[i] = p;
            lox = fmin_(lox, fmin_(p.x, q.x)); loy = fmin_(loy, fmin_(p.y, q.y));
            hix = fmax_(hix, fmax_(p.x, q.x)); hiy = fmax_(hiy, fmax_(p.y, q.y));
            const v2 r = sub(sh.verts[i], sh.localCenter);
            r2 = fmax_(r2, dot(r, r));
        }
        const float d = fabsf(sB.a - sB.a0);
        const float touch = 1.25f * kLinearSlop + 0.002f;   // target + tolerance + safety
        if (d < 0.5f) {
            const float m = touch + sqrtf(r2) * d * d * 0.125f;
            float flox = kFltMax, floy = kFltMax, fhix = -kFltMax, fhiy = -kFltMax;
            for (int i = 0; i < pA.count; i++) {
                flox = fmin_(flox, pA.v[i].x); floy = fmin_(floy, pA.v[i].y);
                fhix = fmax_(fhix, pA.v[i].x); fhiy = fmax_(fhiy, pA.v[i].y);
            }
            far_apart = lox - m > fhix || flox > hix + m || loy - m > fhiy || floy > hiy + m;
            if (far_apart) B2L_STAT(1, 1);
        }
        // Second shortcut, for RESTING and sliding contacts (the bulk of the list: feet on the ground):
        // over the sweep no vertex moves farther than D = |c - c0| + R * |a - a0| from where it starts.
        // If at the start the whole polygon lies on one side of a face line of the fixture by more than
        // D + 1.25 linearSlop (+ safety), it stays beyond the touching distance of that line -- hence of
        // the fixture, which lies on the line's other side -- for the whole step: alpha = 1.  (Resting
        // contacts sit at 3 linearSlop: the position solver leaves one slop of the 2 x polygonRadius.)
        if (!far_apart) {
            const v2 dc = sub(sB.c, sB.c0);
            const float D = sqrtf(dot(dc, dc)) + sqrtf(r2) * d + touch;
            const int nf = pA.count == 2 ? 1 : pA.count;   // an edge has one line, a box four
            for (int e = 0; e < nf && !far_apart; e++) {
                const v2 a0 = pA.v[e], a1 = pA.v[e + 1 < pA.count ? e + 1 : 0];
                v2 n = V(a1.y - a0.y, a0.x - a1.x);          // normal of the line a0-a1 (either sign)
                const float len = sqrtf(dot(n, n));
                if (len < 1e-6f) continue;
                n = scl(1.0f / len, n);
                float smin = kFltMax, smax = -kFltMax;
                for (int i = 0; i < sh.count; i++) {
                    const float sd = dot(n, sub(p0[i], a0));
                    smin = fmin_(smin, sd); smax = fmax_(smax, sd);
                }
                if (pA.count == 2) far_apart = smin > D || smax < -D;   // an edge is two-sided
                else far_apart = smin > D;                               // outside this face of the box
            }
            if (far_apart) B2L_STAT(2, 1);
        }
    }
    return far_apart;
}

// b2Contact::Update of one (body, static fixture) pair during SolveTOI: the manifold at the body's current
// transform, impulses carried over by feature id, the warm-start store kept in step with it, listener events.
// The manifold-table capacity rule of world_step applies.
template <typename Scene>
__device__ __noinline__ bool toi_pair_update(typename Scene::World &W, int b, int f, Manifold &m, float &friction) {
    constexpr int kSlots = Scene::kSlots, kMaxVC = Scene::kMaxVC, NE = Scene::NE, NP = Scene::NP, NF = NE + NP;
    const ShapeConst &sh = Scene::shape(b);
    const uint32_t pair = (uint32_t)(b * NF + f);
    int slot = -1, freeSlot = -1, used = 0;
    for (int s = 0; s < kSlots; s++) {
        if (W.slot_key[s] >> 16) { used++; if ((W.slot_key[s] & 0xffffu) == pair) slot = s; }
        else if (freeSlot < 0) freeSlot = s;
    }
    const bool was = slot >= 0;
    float lox = kFltMax, loy = kFltMax, hix = -kFltMax, hiy = -kFltMax;
    for (int i = 0; i < sh.count; i++) {
        const v2 p = xmul(W.b[b].xf, sh.verts[i]);
        lox = fmin_(lox, p.x); loy = fmin_(loy, p.y); hix = fmax_(hix, p.x); hiy = fmax_(hiy, p.y);
    }
    const float ext = kPolygonRadius + kAabbExtension;
    float efric = 0.0f;
    m.pointCount = 0;
    if (NP > 0 && f >= NE) {
        if constexpr (NP > 0) {
            float x0, ylo, x1, yhi;
            Scene::poly(W, f - NE, x0, ylo, x1, yhi, efric);
            if (!(lox - ext > x1 + ext || x0 - ext > hix + ext || loy - ext > yhi + ext || ylo - ext > hiy + ext)) {
                StaticBox sb;
                static_box(sb, x0, ylo, x1, yhi);
                collide_polygons(m, sb, sh, W.b[b].xf);
            }
        }
    } else {
        v2 v1, v2_;
        Scene::edge(W, f, v1, v2_, efric);
        const float elox = fmin_(v1.x, v2_.x) - ext, ehix = fmax_(v1.x, v2_.x) + ext;
        const float eloy = fmin_(v1.y, v2_.y) - ext, ehiy = fmax_(v1.y, v2_.y) + ext;
        if (!(lox - ext > ehix || elox > hix + ext || loy - ext > ehiy || eloy > hiy + ext))
            collide_edge_polygon(m, v1, v2_, sh, W.b[b].xf);
    }
    friction = sqrtf(efric * sh.friction);
    bool touching = m.pointCount > 0;
    if (touching && !was && (used >= kMaxVC || freeSlot < 0)) {
        touching = false;
        m.pointCount = 0;
        W.flags |= kFlagOverflow;
    }
    if (touching) {
        uint32_t id[2] = {0u, 0u};
        float nI[2] = {0.0f, 0.0f}, tI[2] = {0.0f, 0.0f};
        for (int p = 0; p < m.pointCount; p++) {
            id[p] = m.pts[p].id;
            if (was) {
                const int oc = (int)(W.slot_key[slot] >> 16) - 1;
                for (int q = 0; q < oc; q++)
                    if (W.slot_id[slot][q] == m.pts[p].id) { nI[p] = W.slot_nI[slot][q]; tI[p] = W.slot_tI[slot][q]; break; }
            }
        }
        const int s = was ? slot : freeSlot;
        W.slot_key[s] = pair | ((uint32_t)(m.pointCount + 1) << 16);
        for (int p = 0; p < 2; p++) { W.slot_id[s][p] = id[p]; W.slot_nI[s][p] = nI[p]; W.slot_tI[s][p] = tI[p]; }
    } else if (was) {
        W.slot_key[slot] = 0u;
    }
    if (touching != was) Scene::on_event(W, b, touching);
    return touching;
}

// b2ContactManager::FindNewContacts for one body: pairs whose fat AABBs overlap (the body's AABB over its sweep,
// b2Body::SynchronizeFixtures) join the contact list
template <typename Scene>
__device__ __noinline__ void toi_add_candidates(typename Scene::World &W, ToiCand *cand, int &ncand, float *alphaS, int *statId,
                                                int &nstat, int body) {
    constexpr int NE = Scene::NE, NP = Scene::NP;
    const ShapeConst &sh = Scene::shape(body);
    const Body &B = W.b[body];
    const xform xf0 = sweep_xf(body_sweep(B, sh), 0.0f);
    float lox = kFltMax, loy = kFltMax, hix = -kFltMax, hiy = -kFltMax;
    for (int i = 0; i < sh.count; i++) {
        const v2 p = xmul(xf0, sh.verts[i]), q = xmul(B.xf, sh.verts[i]);
        lox = fmin_(lox, fmin_(p.x, q.x)); loy = fmin_(loy, fmin_(p.y, q.y));
        hix = fmax_(hix, fmax_(p.x, q.x)); hiy = fmax_(hiy, fmax_(p.y, q.y));
    }
    const float ext = kPolygonRadius + kAabbExtension;
    int e_lo, e_hi, p_lo = 0, p_hi = -1;
    Scene::edge_range(W, lox - ext, hix + ext, e_lo, e_hi);
    if constexpr (NP > 0) Scene::poly_range(W, lox - ext, hix + ext, p_lo, p_hi);
    for (int f = (NP > 0 && p_hi >= p_lo) ? NE + p_hi : e_hi; f >= e_lo;) {
        float elox, ehix, eloy, ehiy;
        if (NP > 0 && f >= NE) {
            float x0 = 0.0f, ylo = 0.0f, x1 = 0.0f, yhi = 0.0f, fr;
            if constexpr (NP > 0) Scene::poly(W, f - NE, x0, ylo, x1, yhi, fr);
            elox = x0 - ext; ehix = x1 + ext; eloy = ylo - ext; ehiy = yhi + ext;
        } else {
            v2 v1, v2_;
            float fr;
            Scene::edge(W, f, v1, v2_, fr);
            elox = fmin_(v1.x, v2_.x) - ext; ehix = fmax_(v1.x, v2_.x) + ext;
            eloy = fmin_(v1.y, v2_.y) - ext; ehiy = fmax_(v1.y, v2_.y) + ext;
        }
        const int fcur = f;
        f--;
        if (f > e_hi && f < NE + p_lo) f = e_hi;   // from the lowest candidate polygon down to the highest candidate edge
        if (lox - ext > ehix || elox > hix + ext || loy - ext > ehiy || eloy > hiy + ext) continue;
        bool have = false;
        for (int i = 0; i < ncand; i++) have = have || (cand[i].body == body && cand[i].f == fcur);
        if (have) continue;
        if (ncand >= kMaxToiCand) { W.flags |= kFlagOverflow; continue; }
        ToiCand &c = cand[ncand++];
        c.body = body; c.f = fcur; c.toiCount = 0; c.toiValid = false; c.enabled = true; c.toi = 1.0f;
        const int sid = Scene::kOneStaticBody ? 0 : fcur;
        int si = -1;
        for (int k = 0; k < nstat; k++) if (statId[k] == sid) { si = k; break; }
        if (si < 0) { si = nstat++; statId[si] = sid; alphaS[si] = 0.0f; }
        c.sidx = si;
    }
}

// b2Island::SolveTOI for the island {static bodies, body}: every contact has `body` as its only movable body
template <typename Scene>
__device__ __noinline__ void island_solve_toi(typename Scene::World &W, int body, const Manifold *mf, const float *fric, int nic, float h) {
    constexpr int kMaxVC = Scene::kMaxVC;
    const ShapeConst &sh = Scene::shape(body);
    Body &B = W.b[body];
    const float mB = sh.invMass, iB = sh.invI;
    v2 cB = B.c, vB = B.v;
    float aB = B.a, wB = B.w;
    // SolveTOIPositionConstraints, <= 20 iterations
    for (int it = 0; it < 20; it++) {
        float minSep = 0.0f;
        for (int k = 0; k < nic; k++) {
            const Manifold &m = mf[k];
            for (int p = 0; p < m.pointCount; p++) {
                xform xfB;
                xfB.q = rot_of(aB);
                xfB.p = sub(cB, rmul(xfB.q, sh.localCenter));
                v2 normal, point;
                float separation;
                if (m.type == 0) {
                    normal = m.localNormal;
                    const v2 plane = m.localPoint;
                    const v2 clip = xmul(xfB, m.pts[p].localPoint);
                    separation = dot(sub(clip, plane), normal) - kPolygonRadius - kPolygonRadius;
                    point = clip;
                } else {
                    normal = rmul(xfB.q, m.localNormal);
                    const v2 plane = xmul(xfB, m.localPoint);
                    const v2 clip = m.pts[p].localPoint;
                    separation = dot(sub(clip, plane), normal) - kPolygonRadius - kPolygonRadius;
                    point = clip;
                    normal = neg(normal);
                }
                const v2 rB = sub(point, cB);
                minSep = fmin_(minSep, separation);
                const float C = clampf(kToiBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                const float rnB = crs(rB, normal);
                const float K = mB + iB * rnB * rnB;
                const float impulse = K > 0.0f ? -C / K : 0.0f;
                const v2 P = scl(impulse, normal);
                cB = add(cB, scl(mB, P));
                aB += iB * crs(rB, P);
            }
        }
        if (minSep >= -1.5f * kLinearSlop) break;
    }
    // leap of faith to the new safe state
    B.c0 = cB; B.a0 = aB;
    // InitializeVelocityConstraints (no warm starting: impulses start from zero)
    VC vc[kMaxVC];
    for (int k = 0; k < nic; k++) {
        VC &q = vc[k];
        const Manifold &m = mf[k];
        q.pointCount = m.pointCount; q.friction = fric[k];
        xform xfB;
        xfB.q = rot_of(aB);
        xfB.p = sub(cB, rmul(xfB.q, sh.localCenter));
        xform xfA;
        xfA.p = V(0.0f, 0.0f); xfA.q.s = 0.0f; xfA.q.c = 1.0f;
        v2 pts[2];
        if (m.type == 0) {  // b2WorldManifold::Initialize, e_faceA
            q.normal = rmul(xfA.q, m.localNormal);
            const v2 plane = xmul(xfA, m.localPoint);
            for (int p = 0; p < q.pointCount; p++) {
                const v2 clip = xmul(xfB, m.pts[p].localPoint);
                const v2 cA = add(clip, scl(kPolygonRadius - dot(sub(clip, plane), q.normal), q.normal));
                const v2 cBp = sub(clip, scl(kPolygonRadius, q.normal));
                pts[p] = scl(0.5f, add(cA, cBp));
            }
        } else {            // e_faceB
            const v2 nrm = rmul(xfB.q, m.localNormal);
            const v2 plane = xmul(xfB, m.localPoint);
            for (int p = 0; p < q.pointCount; p++) {
                const v2 clip = xmul(xfA, m.pts[p].localPoint);
                const v2 cBp = add(clip, scl(kPolygonRadius - dot(sub(clip, plane), nrm), nrm));
                const v2 cA = sub(clip, scl(kPolygonRadius, nrm));
                pts[p] = scl(0.5f, add(cA, cBp));
            }
            q.normal = neg(nrm);
        }
        for (int p = 0; p < q.pointCount; p++) {
            VCP &cp = q.p[p];
            cp.nI = 0.0f; cp.tI = 0.0f;
            cp.rB = sub(pts[p], cB);
            const float rnB = crs(cp.rB, q.normal);
            const float kN = mB + iB * rnB * rnB;
            cp.normalMass = kN > 0.0f ? 1.0f / kN : 0.0f;
            const v2 tangent = crs_vs(q.normal, 1.0f);
            const float rtB = crs(cp.rB, tangent);
            const float kT = mB + iB * rtB * rtB;
            cp.tangentMass = kT > 0.0f ? 1.0f / kT : 0.0f;
        }
        if (q.pointCount == 2) {
            const float rn1B = crs(q.p[0].rB, q.normal), rn2B = crs(q.p[1].rB, q.normal);
            const float k11 = mB + iB * rn1B * rn1B, k22 = mB + iB * rn2B * rn2B, k12 = mB + iB * rn1B * rn2B;
            if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
                q.k11 = k11; q.k12 = k12; q.k22 = k22;
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = 1.0f / det;
                q.n11 = det * k22; q.n12 = -det * k12; q.n22 = det * k11;
            } else q.pointCount = 1;
        }
    }
    // SolveVelocityConstraints x 180.  The iteration is a deterministic map of (v, w, accumulated impulses): as soon
    // as an iteration reproduces the state of the previous one (a fixed point) or of the one before that (a 2-cycle
    // in the last bits) the remaining iterations are known without running them -- the state stays, or alternates,
    // and the result after the 180th iteration is picked by parity.  Bit-exact by construction (the comparison is on
    // the bit patterns); a single body against static geometry gets there within 10-20 iterations in ~90 % of the
    // TOI sub-steps (measured on the CPU checker), which matters because a warp waits for its slowest lane.
    uint32_t hist[2][3 + 4 * kMaxVC];   // state after the previous two iterations: slot (it & 1)
    float raw[kMaxVC][2];               // the raw tangent increment of every point in the current iteration
#pragma unroll 1
    for (int it = 0; it < 180; it++) {
        for (int k = 0; k < nic; k++) {
            VC &q = vc[k];
            const v2 normal = q.normal, tangent = crs_vs(normal, 1.0f);
            for (int p = 0; p < q.pointCount; p++) {
                VCP &cp = q.p[p];
                const v2 dv = add(vB, crs_sv(wB, cp.rB));
                const float vt = dot(dv, tangent) - 0.0f;
                float lambda = cp.tangentMass * (-vt);
                raw[k][p] = lambda;                       // for the friction-drift exit below
                const float maxF = q.friction * cp.nI;
                const float newImp = clampf(cp.tI + lambda, -maxF, maxF);
                lambda = newImp - cp.tI;
                cp.tI = newImp;
                const v2 P = scl(lambda, tangent);
                vB = add(vB, scl(mB, P)); wB += iB * crs(cp.rB, P);
            }
            if (q.pointCount == 1) {
                VCP &cp = q.p[0];
                const v2 dv = add(vB, crs_sv(wB, cp.rB));
                const float vn = dot(dv, normal);
                float lambda = -cp.normalMass * (vn - 0.0f);
                const float newImp = fmax_(cp.nI + lambda, 0.0f);
                lambda = newImp - cp.nI;
                cp.nI = newImp;
                const v2 P = scl(lambda, normal);
                vB = add(vB, scl(mB, P)); wB += iB * crs(cp.rB, P);
            } else {
                VCP &c1 = q.p[0], &c2 = q.p[1];
                const v2 a = V(c1.nI, c2.nI);
                const v2 dv1 = add(vB, crs_sv(wB, c1.rB)), dv2 = add(vB, crs_sv(wB, c2.rB));
                float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
                v2 b = V(vn1 - 0.0f, vn2 - 0.0f);
                b = sub(b, V(q.k11 * a.x + q.k12 * a.y, q.k12 * a.x + q.k22 * a.y));
                v2 x;
                bool solved = false;
                x = V(-(q.n11 * b.x + q.n12 * b.y), -(q.n12 * b.x + q.n22 * b.y));
                if (x.x >= 0.0f && x.y >= 0.0f) solved = true;
                if (!solved) {
                    x.x = -c1.normalMass * b.x; x.y = 0.0f;
                    vn2 = q.k12 * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) solved = true;
                }
                if (!solved) {
                    x.x = 0.0f; x.y = -c2.normalMass * b.y;
                    vn1 = q.k12 * x.y + b.x;
                    if (x.y >= 0.0f && vn1 >= 0.0f) solved = true;
                }
                if (!solved) {
                    x.x = 0.0f; x.y = 0.0f;
                    if (b.x >= 0.0f && b.y >= 0.0f) solved = true;
                }
                if (solved) {
                    const v2 d = sub(x, a);
                    const v2 P1 = scl(d.x, normal), P2 = scl(d.y, normal);
                    vB = add(vB, scl(mB, add(P1, P2)));
                    wB += iB * (crs(c1.rB, P1) + crs(c2.rB, P2));
                    c1.nI = x.x; c2.nI = x.y;
                }
            }
        }
        B2L_STAT(5, 1);
        {   // periodicity check
            uint32_t cur[3 + 4 * kMaxVC];
            cur[0] = __float_as_uint(vB.x); cur[1] = __float_as_uint(vB.y); cur[2] = __float_as_uint(wB);
            int nw = 3;
            for (int k = 0; k < nic; k++)
                for (int p = 0; p < vc[k].pointCount; p++) { cur[nw++] = __float_as_uint(vc[k].p[p].nI); cur[nw++] = __float_as_uint(vc[k].p[p].tI); }
            bool same1 = it >= 1, same2 = it >= 2;
            const uint32_t *h1 = hist[(it - 1) & 1], *h2 = hist[it & 1];
            for (int q = 0; q < nw; q++) { same1 = same1 && cur[q] == h1[q]; same2 = same2 && cur[q] == h2[q]; }
            if (same1) break;
            if (same2) {
                if ((179 - it) & 1) { vB.x = __uint_as_float(h1[0]); vB.y = __uint_as_float(h1[1]); wB = __uint_as_float(h1[2]); }
                break;
            }
            // Third exit, "friction drift": v, w and every normal impulse are back where the iteration started, but a
            // tangent impulse keeps creeping by a constant, tiny amount per iteration (its effect on v is undone by the
            // normal step, or rounds away), so the state never repeats -- 16 % of LunarLander's TOI sub-steps, 24 % of
            // BipedalWalker's.  Only (v, w) leave this loop (TOI impulses are discarded), so it may stop once the
            // remaining iterations are KNOWN to reproduce this one.  The sub-steps of an iteration read tI only in
            // the friction clamp, newImp = clamp(tI + r), and use d = newImp - tI; if every point's d is the same
            // next time, every sub-step gets this iteration's inputs and returns its outputs, by induction over the
            // sub-steps and then over the iterations.  Per point, with r its raw increment of THIS iteration:
            //   (1) the next friction step applies the same increment: clamp(fl(tI + r)) - tI == d;
            //   (2) so do the later ones: tI walks through multiples of its ulp and fl(tI + r) rounds r to the same
            //       multiple as long as tI and tI + r stay inside one binade -- checked for 180 more steps, with
            //       margin -- and r is not exactly half-way between two multiples (ties-to-even would alternate);
            //   (3) tI stays strictly inside the friction cone mu * nI for those steps, so the clamp stays inactive.
            if (it >= 1) {
                bool fixed = cur[0] == h1[0] && cur[1] == h1[1] && cur[2] == h1[2];
                for (int q = 3; q < nw && fixed; q += 2) fixed = cur[q] == h1[q];          // normal impulses
                if (fixed) {
                    bool repeats = true;
                    int q = 3;
                    for (int k = 0; k < nic && repeats; k++) {
                        for (int p = 0; p < vc[k].pointCount && repeats; p++, q += 2) {
                            const VCP &cp = vc[k].p[p];
                            const float tI = cp.tI, d = tI - __uint_as_float(h1[q + 1]), r = raw[k][p];
                            const float maxF = vc[k].friction * cp.nI;
                            const float next = clampf(tI + r, -maxF, maxF);
                            if (d == 0.0f) { repeats = (next - tI) == 0.0f; continue; }      // this point is at rest
                            const uint32_t eb = __float_as_uint(tI) & 0x7f800000u;
                            const float lo2 = __uint_as_float(eb);                           // 2^e <= |tI| < 2^(e+1)
                            const float ulp = lo2 * 1.1920929e-07f;
                            const float m = r / ulp, fr = m - floorf(m);
                            const float end = fabsf(tI + 181.0f * d + r);
                            repeats = eb != 0u && eb != 0x7f800000u && (next - tI) == d && fr != 0.5f &&
                                      fabsf(tI) > 1.0001f * lo2 && end > 1.0001f * lo2 && end < 1.9999f * lo2 &&
                                      end < 0.9999f * maxF && fabsf(tI) < 0.9999f * maxF;
                        }
                    }
                    if (repeats) break;
                }
            }
            for (int q = 0; q < nw; q++) hist[it & 1][q] = cur[q];
        }
    }
    // the TOI impulses are NOT stored for warm starting; integrate positions over the rest of the step
    {
        const v2 tr = scl(h, vB);
        if (dot(tr, tr) > kMaxTranslation * kMaxTranslation) { const float ratio = kMaxTranslation / sqrtf(dot(tr, tr)); vB = scl(ratio, vB); }
        const float rotn = h * wB;
        if (rotn * rotn > kMaxRotation * kMaxRotation) { const float ratio = kMaxRotation / fabsf(rotn); wB *= ratio; }
        cB = add(cB, scl(h, vB));
        aB = aB + h * wB;
    }
    B.c = cB; B.a = aB; B.v = vB; B.w = wB;
    sync_xf(B, sh);
}

// Is there any pair on the step's contact list for which b2TimeOfImpact has to be evaluated at all?  (Same list,
// same sweeps and the same shortcut as the first round of solve_toi; when this returns false solve_toi would find
// every alpha = 1 and change nothing.)  The step kernels use it to finish the common case -- free flight, resting
// contacts -- right away and to hand only the envs with a possible event to the TOI kernel.
template <typename Scene>
__device__ __noinline__ bool toi_needed(typename Scene::World &W) {
    constexpr int NB = Scene::NB;
    ToiCand cand[kMaxToiCand];
    float alphaS[kMaxToiCand];
    int statId[kMaxToiCand];
    int ncand = 0, nstat = 0;
    for (int i = 0; i < NB; i++) W.b[i].alpha0 = 0.0f;
    for (int oi = 0; oi < NB; oi++) toi_add_candidates<Scene>(W, cand, ncand, alphaS, statId, nstat, Scene::body_order(oi));
    for (int ci = 0; ci < ncand; ci++) {
        const ShapeConst &sh = Scene::shape(cand[ci].body);
        const Body &B = W.b[cand[ci].body];
        DProxy pA;
        fixture_proxy<Scene>(W, cand[ci].f, pA);
        if (!toi_cannot_touch(sh, B.xf, body_sweep(B, sh), pA)) return true;
    }
    return false;
}

// `live`: the lanes of this warp that run solve_toi together (0: the caller is on a divergent path, e.g. the step
// embedded in reset(): no warp-level synchronisation).  TOI events are rare per env but a warp of 32 envs has one
// in most steps, and an event costs a 180-iteration solve: the loop below runs in warp-synchronous ROUNDS -- every
// lane looks for its earliest event, then all lanes that found one advance / update / solve at the same time -- so
// a warp pays for the lane with the most events, not for the sum over its lanes.  Per env the sequence of
// operations is exactly the serial b2World::SolveTOI loop.
template <typename Scene>
__device__ __noinline__ void solve_toi(typename Scene::World &W, float dt, bool active, unsigned live) {
    constexpr int NB = Scene::NB, NE = Scene::NE, kMaxVC = Scene::kMaxVC;
    const bool sync = live != 0u;
    ToiCand cand[kMaxToiCand];
    float alphaS[kMaxToiCand];
    int statId[kMaxToiCand];
    int ncand = 0, nstat = 0;
    if (active) {
        for (int i = 0; i < NB; i++) W.b[i].alpha0 = 0.0f;
        for (int oi = 0; oi < NB; oi++) toi_add_candidates<Scene>(W, cand, ncand, alphaS, statId, nstat, Scene::body_order(oi));
    }
    for (;;) {
        // The search for the earliest event, in three passes so that the lanes of a warp run the expensive part
        // together: (1) in contact-list order, the bookkeeping b2World::SolveTOI does before each b2TimeOfImpact (which
        // sweep interval the pair is on, advancing a sweep to it) and the shortcut; pairs that need the real thing
        // go onto a work list with a snapshot of their inputs -- b2TimeOfImpact reads state but changes none, so its
        // evaluations can be taken out of the loop; (2) the w-th work item of every lane, all lanes at once;
        // (3) the minimum, in contact-list order like the serial loop (the first of equal alphas wins).
        int nwork = 0;
        int work_ci[kMaxToiCand];
        float work_alpha0[kMaxToiCand];
        Sweep work_sB[kMaxToiCand];
        for (int ci = 0; ci < ncand; ci++) {
            ToiCand &c = cand[ci];
            if (!c.enabled) continue;
            if (c.toiCount > kMaxSubSteps) continue;
            if (c.toiValid) continue;
            Body &B = W.b[c.body];
            const ShapeConst &sh = Scene::shape(c.body);
            float alpha0 = alphaS[c.sidx];
            if (alphaS[c.sidx] < B.alpha0) { alpha0 = B.alpha0; alphaS[c.sidx] = alpha0; }
            else if (B.alpha0 < alphaS[c.sidx]) {
                alpha0 = alphaS[c.sidx];
                Sweep s = body_sweep(B, sh);
                sweep_advance(s, alpha0);
                B.c0 = s.c0; B.a0 = s.a0; B.alpha0 = s.alpha0;
            }
            DProxy pA;
            fixture_proxy<Scene>(W, c.f, pA);
            const Sweep sB = body_sweep(B, sh);
            B2L_STAT(0, 1);
            if (toi_cannot_touch(sh, B.xf, sB, pA)) { c.toi = 1.0f; c.toiValid = true; }
            else { work_ci[nwork] = ci; work_alpha0[nwork] = alpha0; work_sB[nwork] = sB; nwork++; }
        }
        for (int w = 0; sync ? (__any_sync(live, w < nwork) != 0) : (w < nwork); w++) {
            if (w < nwork) {
                ToiCand &c = cand[work_ci[w]];
                const ShapeConst &sh = Scene::shape(c.body);
                const float alpha0 = work_alpha0[w];
                DProxy pA, pB;
                fixture_proxy<Scene>(W, c.f, pA);
                pB.count = sh.count;
                for (int i = 0; i < sh.count; i++) pB.v[i] = sh.verts[i];
                Sweep sA;
                sA.localCenter = V(0.0f, 0.0f); sA.c0 = V(0.0f, 0.0f); sA.c = V(0.0f, 0.0f); sA.a0 = 0.0f; sA.a = 0.0f; sA.alpha0 = alpha0;
                float beta;
                B2L_STAT(3, 1);
                const int state = time_of_impact(beta, pA, sA, pB, work_sB[w], 1.0f);
                c.toi = state == TOI_TOUCHING ? fmin_(alpha0 + (1.0f - alpha0) * beta, 1.0f) : 1.0f;
                c.toiValid = true;
            }
        }
        int minIdx = -1;
        float minAlpha = 1.0f;
        for (int ci = 0; ci < ncand; ci++) {
            const ToiCand &c = cand[ci];
            if (!c.enabled) continue;
            if (c.toiCount > kMaxSubSteps) continue;
            if (c.toi < minAlpha) { minIdx = ci; minAlpha = c.toi; }
        }
        const bool have = minIdx >= 0 && !(1.0f - 10.0f * kEpsilon < minAlpha);
        if (!(sync ? (__any_sync(live, have) != 0) : have)) break;
        Manifold mf[kMaxVC];
        float fric[kMaxVC];
        int nic = 0, moved = -1;
        float h = 0.0f;
        bool solve = false;
        if (have) {
            ToiCand &mc = cand[minIdx];
            Body &B = W.b[mc.body];
            const ShapeConst &sh = Scene::shape(mc.body);
            // advance the bodies to the TOI
            const float backupS = alphaS[mc.sidx];
            const v2 bc0 = B.c0, bc = B.c;
            const float ba0 = B.a0, ba = B.a, balpha0 = B.alpha0;
            alphaS[mc.sidx] = minAlpha;
            {   // b2Body::Advance
                Sweep s = body_sweep(B, sh);
                sweep_advance(s, minAlpha);
                B.c0 = s.c0; B.a0 = s.a0; B.alpha0 = s.alpha0;
                B.c = B.c0; B.a = B.a0;
                sync_xf(B, sh);
            }
            const bool touching = toi_pair_update<Scene>(W, mc.body, mc.f, mf[0], fric[0]);
            mc.toiValid = false;
            ++mc.toiCount;
            if (!touching) {
                mc.enabled = false;
                alphaS[mc.sidx] = backupS;
                B.c0 = bc0; B.c = bc; B.a0 = ba0; B.a = ba; B.alpha0 = balpha0;
                sync_xf(B, sh);
            } else {
                // the island: the TOI contact, then the body's other touching contacts against static bodies
                nic = 1;
                uint64_t inIsland = 1ull << mc.sidx;   // static bodies (by sidx) already in the island
                for (int ci = 0; ci < ncand; ci++) {
                    ToiCand &oc = cand[ci];
                    if (ci == minIdx || oc.body != mc.body) continue;
                    if (nic == kMaxVC) break;
                    const float backup = alphaS[oc.sidx];
                    if (!((inIsland >> oc.sidx) & 1ull)) alphaS[oc.sidx] = minAlpha;
                    const bool t2 = toi_pair_update<Scene>(W, oc.body, oc.f, mf[nic], fric[nic]);
                    oc.enabled = true;   // b2Contact::Update re-enables the contact
                    if (!t2) { alphaS[oc.sidx] = backup; continue; }
                    nic++;
                    inIsland |= 1ull << oc.sidx;
                }
                solve = true;
                moved = mc.body;
                h = (1.0f - minAlpha) * dt;
            }
        }
        if (sync) __syncwarp(live);
        if (solve) { B2L_STAT(4, 1); island_solve_toi<Scene>(W, moved, mf, fric, nic, h); }
        if (solve) {
            // invalidate all contact TOIs on the displaced body; its moved proxy may create new contacts
            for (int ci = 0; ci < ncand; ci++) if (cand[ci].body == moved) cand[ci].toiValid = false;
            toi_add_candidates<Scene>(W, cand, ncand, alphaS, statId, nstat, moved);
        }
    }
}
