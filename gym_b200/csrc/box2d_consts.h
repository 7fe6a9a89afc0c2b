// box2d_consts.h -- host-side construction of the Box2D-task constants (plain C++, no CUDA runtime).
//
// The polygon / mass / joint constants of the LunarLander and BipedalWalker scenes are evaluated once on the host,
// with the float32 operation sequence Box2D uses, and handed to the device as `__constant__` data (b200gym.cu) --
// or to the host build of the same solver headers that tests/hostsim compiles to check them on a CPU.
#pragma once
#include <cmath>
#include <cstring>

#include "lunar.cuh"
#include "walker.cuh"

namespace b2l_host {

// Shape / mass constants, evaluated on the host with the float32 operation sequence of
// b2PolygonShape::Set / ComputeCentroid / ComputeMass and b2Body::ResetMassData.
inline void shape(b2l::ShapeConst &sh, const b2l::v2 *hull, int n, float density, float friction, bool box) {
    using b2l::v2;
    auto V = [](float x, float y) { v2 r; r.x = x; r.y = y; return r; };
    sh.count = n;
    sh.friction = friction;
    for (int i = 0; i < n; i++) sh.verts[i] = hull[i];
    for (int i = 0; i < n; i++) {
        const int i2 = i + 1 < n ? i + 1 : 0;
        const float ex = sh.verts[i2].x - sh.verts[i].x, ey = sh.verts[i2].y - sh.verts[i].y;
        const float nx = 1.0f * ey, ny = -1.0f * ex;
        const float len = sqrtf(nx * nx + ny * ny), inv = 1.0f / len;
        sh.normals[i] = V(inv * nx, inv * ny);
    }
    {   // ComputeCentroid
        float cx = 0.0f, cy = 0.0f, area = 0.0f;
        const float inv3 = 1.0f / 3.0f;
        for (int i = 0; i < n; i++) {
            const v2 p1 = V(0.0f, 0.0f), p2 = sh.verts[i], p3 = i + 1 < n ? sh.verts[i + 1] : sh.verts[0];
            const float e1x = p2.x - p1.x, e1y = p2.y - p1.y, e2x = p3.x - p1.x, e2y = p3.y - p1.y;
            const float D = e1x * e2y - e1y * e2x, tri = 0.5f * D;
            area += tri;
            const float w = tri * inv3;
            cx = cx + w * ((p1.x + p2.x) + p3.x);
            cy = cy + w * ((p1.y + p2.y) + p3.y);
        }
        const float ia = 1.0f / area;
        sh.centroid = V(ia * cx, ia * cy);
    }
    if (box) {  // b2PolygonShape::SetAsBox writes exact normals and a zero centroid
        sh.normals[0] = V(0.0f, -1.0f); sh.normals[1] = V(1.0f, 0.0f); sh.normals[2] = V(0.0f, 1.0f); sh.normals[3] = V(-1.0f, 0.0f);
        sh.centroid = V(0.0f, 0.0f);
    }
    // ComputeMass
    float sx = 0.0f, sy = 0.0f;
    for (int i = 0; i < n; i++) { sx = sx + sh.verts[i].x; sy = sy + sh.verts[i].y; }
    const float invn = 1.0f / (float)n;
    sx = invn * sx; sy = invn * sy;
    float cx = 0.0f, cy = 0.0f, area = 0.0f, I = 0.0f;
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < n; i++) {
        const float e1x = sh.verts[i].x - sx, e1y = sh.verts[i].y - sy;
        const v2 nxt = i + 1 < n ? sh.verts[i + 1] : sh.verts[0];
        const float e2x = nxt.x - sx, e2y = nxt.y - sy;
        const float D = e1x * e2y - e1y * e2x, tri = 0.5f * D;
        area += tri;
        const float w = tri * inv3;
        cx = cx + w * (e1x + e2x);
        cy = cy + w * (e1y + e2y);
        const float intx2 = e1x * e1x + e2x * e1x + e2x * e2x, inty2 = e1y * e1y + e2y * e1y + e2y * e2y;
        I += (0.25f * inv3 * D) * (intx2 + inty2);
    }
    const float mass = density * area;
    const float ia = 1.0f / area;
    cx = ia * cx; cy = ia * cy;
    const float mcx = cx + sx, mcy = cy + sy;
    float mI = density * I;
    mI += mass * ((mcx * mcx + mcy * mcy) - (cx * cx + cy * cy));
    sh.invMass = 1.0f / mass;
    const float lcx = sh.invMass * (mass * mcx), lcy = sh.invMass * (mass * mcy);
    const float bI = mI - mass * (lcx * lcx + lcy * lcy);
    sh.invI = 1.0f / bI;
    sh.localCenter = V(lcx, lcy);
}

// LunarLander scene: lunar_lander.py:327,354-412
inline void lunar_consts(lunar::Consts &c) {
    using b2l::v2;
    auto V = [](float x, float y) { v2 r; r.x = x; r.y = y; return r; };
    memset(&c, 0, sizeof c);
    const double SCALE = 30.0;
    const double LP[6][2] = {{17, -10}, {17, 0}, {14, 17}, {-14, 17}, {-17, 0}, {-17, -10}};  // hull order of LANDER_POLY
    v2 hull[6];
    for (int i = 0; i < 6; i++) hull[i] = V((float)(LP[i][0] / SCALE), (float)(LP[i][1] / SCALE));
    shape(c.shape[0], hull, 6, 5.0f, 0.1f, false);                       // lunar_lander.py:354-368
    const float hx = (float)(2 / SCALE), hy = (float)(8 / SCALE);              // LEG_W, LEG_H
    const v2 box[4] = {V(-hx, -hy), V(hx, -hy), V(hx, hy), V(-hx, hy)};
    shape(c.shape[1], box, 4, 1.0f, 0.2f, true);                         // :379-392
    for (int li = 0; li < 2; li++) {
        const int i = li == 0 ? -1 : +1;
        c.jd[li].bodyA = 0;
        c.jd[li].bodyB = 1 + li;
        c.jd[li].anchorA = V(0.0f, 0.0f);                                      // :396
        c.jd[li].anchorB = V((float)(i * 20 / SCALE), (float)(18 / SCALE));   // :397
        c.motorSpeed[li] = (float)(+0.3 * i);                                  // :401
        if (i == -1) { c.jd[li].lower = (float)(+0.9 - 0.5); c.jd[li].upper = (float)(+0.9); }   // :403-410
        else { c.jd[li].lower = (float)(-0.9); c.jd[li].upper = (float)(-0.9 + 0.5); }
        c.leg_x0[li] = (float)(600 / SCALE / 2 - i * 20 / SCALE);             // :381
        c.leg_a0[li] = (float)(i * 0.05);                                      // :382
    }
    const double Wd = 600 / SCALE;
    for (int e = 0; e < 11; e++) c.chunk_x[e] = (float)(Wd / (11 - 1) * e);    // :327
    c.world_w = (float)Wd;
    c.lander_x0 = (float)(600 / SCALE / 2);
    c.y0 = (float)(400 / SCALE);
}

// BipedalWalker scene: bipedal_walker.py:55-78,442-500
inline void walker_consts(walker::Consts &c) {
    using b2l::v2;
    auto V = [](float x, float y) { v2 r; r.x = x; r.y = y; return r; };
    memset(&c, 0, sizeof c);
    const double SCALE = 30.0, LEG_DOWN = -8 / SCALE, LEG_W = 8 / SCALE, LEG_H = 34 / SCALE;
    const double TERRAIN_STEP = 14 / SCALE, TERRAIN_HEIGHT = 400 / SCALE / 4;
    const double HP[5][2] = {{34, -8}, {34, 1}, {6, 9}, {-30, 9}, {-30, -8}};  // hull order of HULL_POLY
    v2 hull[5];
    for (int i = 0; i < 5; i++) hull[i] = V((float)(HP[i][0] / SCALE), (float)(HP[i][1] / SCALE));
    shape(c.shape[0], hull, 5, 5.0f, 0.1f, false);                         // HULL_FD, bipedal_walker.py:55-62
    {
        const float hx = (float)(LEG_W / 2), hy = (float)(LEG_H / 2);
        const v2 box[4] = {V(-hx, -hy), V(hx, -hy), V(hx, hy), V(-hx, hy)};
        shape(c.shape[1], box, 4, 1.0f, 0.2f, true);                       // LEG_FD :64-70
    }
    {
        const float hx = (float)(0.8 * LEG_W / 2), hy = (float)(LEG_H / 2);
        const v2 box[4] = {V(-hx, -hy), V(hx, -hy), V(hx, hy), V(-hx, hy)};
        shape(c.shape[2], box, 4, 1.0f, 0.2f, true);                       // LOWER_FD :72-78
    }
    for (int li = 0; li < 2; li++) {
        const int i = li == 0 ? -1 : +1;
        b2l::JointDef &hip = c.jd[2 * li], &knee = c.jd[2 * li + 1];
        hip.bodyA = 0; hip.bodyB = 1 + 2 * li;                                 // :465-476
        hip.anchorA = V(0.0f, (float)LEG_DOWN); hip.anchorB = V(0.0f, (float)(LEG_H / 2));
        hip.lower = -0.8f; hip.upper = 1.1f;
        knee.bodyA = 1 + 2 * li; knee.bodyB = 2 + 2 * li;                      // :487-498
        knee.anchorA = V(0.0f, (float)(-LEG_H / 2)); knee.anchorB = V(0.0f, (float)(LEG_H / 2));
        knee.lower = -1.6f; knee.upper = -0.1f;
        c.leg_a0[li] = (float)(i * 0.05);
    }
    const double init_x = TERRAIN_STEP * 20 / 2, init_y = TERRAIN_HEIGHT + 2 * LEG_H;   // :442-443
    c.init_x = (float)init_x;
    c.init_y = (float)init_y;
    c.leg_y = (float)(init_y - LEG_H / 2 - LEG_DOWN);
    c.lower_y = (float)(init_y - LEG_H * 3 / 2 - LEG_DOWN);
}

}  // namespace b2l_host
