// ORACLE (test infrastructure, NOT product code): scalar software restatement of the reference's render path
//   V4R vertex stage          src/3rdparty/v4r/src/pipelines/shaders/uber.vert:53-110
//   V4R fragment stage         src/3rdparty/v4r/src/pipelines/shaders/uber.frag:112-141 (the non-V4R_BLINN_PHONG branch)
//   projection                 src/3rdparty/v4r/src/v4r.cpp:35-45 (hfov 100 deg, near 0.01, far 120; env_renderer.hpp:34-38)
//   raster / depth state       src/3rdparty/v4r/src/vulkan_state.cpp:588-606 (cull back, CCW front, LESS_OR_EQUAL, D32)
//   light + materials          src/libs/v4r_rendering/src/v4r_env_renderer.cpp:204-220 (light (0,4,2) colour 0.66, shininess 300)
//   output layout              RGBA8 rows top-down, view index = env*A + agent (v4r_env_renderer.cpp:357-361)
// The fixed-function part (clip, 8-bit sub-pixel snap, top-left rule, perspective-correct varyings) follows the
// Vulkan specification's rasterisation rules; the reference's actual GPU is not observable here (PARITY UNPINNED
// for assembled frames: the north star allows +-1 LSB per channel against real V4R frames).  PINNED: shadeFragment against the
// shader's own compute_color() text compiled with glm, normalMatrix against the vertex stage's glm expression, the view matrices
// against Magnum's Camera feature, the model matrices against the reference's scene graph (tests/test_ref_shim.py).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "orc_env.hpp"

namespace orc {

#include "orc_meshes.inc"

struct MeshRef { const uint32_t (*vtx)[6]; int nv; const uint16_t *idx; int ni; };
inline MeshRef meshRef(int type) {
    switch (type) {
        case MESH_BOX: return {mesh_box_vtx, 24, mesh_box_idx, 36};
        case MESH_CAPSULE: return {mesh_capsule_vtx, 66, mesh_capsule_idx, 384};
        case MESH_SPHERE: return {mesh_sphere_vtx, 42, mesh_sphere_idx, 240};
        case MESH_CONE: return {mesh_cone_vtx, 12, mesh_cone_idx, 36};
        default: return {mesh_cylinder_vtx, 26, mesh_cylinder_idx, 72};
    }
}
inline float bitsToFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

struct Projection {
    float p00, p11, p22, p32;
    Projection(int W, int H, float hfovDeg = 100.0f, float nearZ = 0.01f, float farZ = 120.0f) {
        const float aspect = float(W) / float(H);
        const float halfTan = crtan((hfovDeg * 0.01745329251994329576923690768489f) / 2.0f);  // glm::radians(hfov)/2
        p00 = 1.0f / halfTan;
        p11 = -aspect / halfTan;
        p22 = farZ / (nearZ - farZ);
        p32 = farZ * nearZ / (nearZ - farZ);
    }
};

struct ClipVert {
    float cx, cy, cz, cw;  // clip space
    float px, py, pz;      // camera space position (varying)
    float nx, ny, nz;      // camera space normal (varying)
};
inline ClipVert lerpVert(const ClipVert &a, const ClipVert &b, float t) {
    ClipVert o;
    const float *pa = &a.cx, *pb = &b.cx;
    float *po = &o.cx;
    for (int i = 0; i < 10; ++i) po[i] = pa[i] + t * (pb[i] - pa[i]);
    return o;
}

struct SetupTri {  // a screen-space triangle ready for per-pixel evaluation
    int32_t x[3], y[3];  // snapped to 1/256 pixel
    float z[3], rw[3];   // ndc depth, 1/w
    float p[3][3], n[3][3];
    int64_t area;  // > 0 (front-facing, sign flipped)
    int color;
};

inline int32_t snap(float v) { return int32_t(floorf(v * 256.0f + 0.5f)); }

// clip one triangle against z >= 0 and z <= w, project, snap, cull; append 0..3 triangles
inline void clipAndSetup(const ClipVert tri[3], int W, int H, int color, std::vector<SetupTri> &out) {
    ClipVert poly[8], tmp[8];
    int n = 3;
    for (int i = 0; i < 3; ++i) poly[i] = tri[i];
    for (int plane = 0; plane < 2; ++plane) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const ClipVert &a = poly[i], &b = poly[(i + 1) % n];
            const float da = plane == 0 ? a.cz : a.cw - a.cz;
            const float db = plane == 0 ? b.cz : b.cw - b.cz;
            const bool ina = da >= 0.0f, inb = db >= 0.0f;
            if (ina) tmp[m++] = a;
            if (ina != inb) {
                // always interpolate from the inside vertex so that shared edges clip identically
                if (ina) tmp[m++] = lerpVert(a, b, da / (da - db));
                else tmp[m++] = lerpVert(b, a, db / (db - da));
            }
        }
        n = m;
        for (int i = 0; i < n; ++i) poly[i] = tmp[i];
        if (n < 3) return;
    }
    const float hw = float(W) * 0.5f, hh = float(H) * 0.5f;
    int32_t sx[8], sy[8];
    float sz[8], rw[8];
    for (int i = 0; i < n; ++i) {
        const float r = 1.0f / poly[i].cw;
        rw[i] = r;
        sx[i] = snap((poly[i].cx * r) * hw + hw);
        sy[i] = snap((poly[i].cy * r) * hh + hh);
        sz[i] = poly[i].cz * r;
    }
    for (int k = 1; k + 1 < n; ++k) {
        const int id[3] = {0, k, k + 1};
        SetupTri t;
        for (int j = 0; j < 3; ++j) {
            const int v = id[j];
            t.x[j] = sx[v]; t.y[j] = sy[v]; t.z[j] = sz[v]; t.rw[j] = rw[v];
            t.p[j][0] = poly[v].px; t.p[j][1] = poly[v].py; t.p[j][2] = poly[v].pz;
            t.n[j][0] = poly[v].nx; t.n[j][1] = poly[v].ny; t.n[j][2] = poly[v].nz;
        }
        const int64_t area2 = int64_t(t.x[1] - t.x[0]) * int64_t(t.y[2] - t.y[0]) - int64_t(t.y[1] - t.y[0]) * int64_t(t.x[2] - t.x[0]);
        if (area2 >= 0) continue;  // back-facing (visually clockwise with y down) or degenerate
        t.area = -area2;
        t.color = color;
        out.push_back(t);
    }
}

inline float pow300(float x) {
    const float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8, x32 = x16 * x16, x64 = x32 * x32, x128 = x64 * x64, x256 = x128 * x128;
    return ((x256 * x32) * x8) * x4;
}
inline uint8_t toUnorm8(float c) {
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    return uint8_t(floorf(c * 255.0f + 0.5f));
}

// The lit fragment stage, uber.frag:112-141 (non-V4R_BLINN_PHONG branch) with the one light the environment sets up: camera-space
// position (0,4,2), colour 0.66 (v4r_env_renderer.cpp:220), specular (1,1,1), shininess 300 (:204-213).  P = interpolated camera-space
// position, N = interpolated (unnormalised) normal, diffuse = material colour.  Pinned against the shader text itself compiled with
// glm (oracle/ref_shim/shade_shim.cpp) by tests/test_ref_shim.py.
inline void shadeFragment(const float P[3], const float N[3], const float diffuse[3], float Lo[3]) {
    const float cd[3] = {-P[0], -P[1], -P[2]};
    const float ld[3] = {0.0f + cd[0], 4.0f + cd[1], 2.0f + cd[2]};
    const float ldi = 1.0f / sqrtf((ld[0] * ld[0] + ld[1] * ld[1]) + ld[2] * ld[2]);
    const float nl[3] = {ld[0] * ldi, ld[1] * ldi, ld[2] * ldi};
    const float nni = 1.0f / sqrtf((N[0] * N[0] + N[1] * N[1]) + N[2] * N[2]);
    const float nn[3] = {N[0] * nni, N[1] * nni, N[2] * nni};
    const float ndl = (nn[0] * nl[0] + nn[1] * nl[1]) + nn[2] * nl[2];
    const float intensity = ndl > 0.0f ? ndl : 0.0f;
    float spec = 0.0f;
    if (intensity > 0.001f) {
        // reflect(I, N) = I - 2*dot(N, I)*N with I = -nl
        const float dni = -ndl;
        const float refl[3] = {-nl[0] - (2.0f * dni) * nn[0], -nl[1] - (2.0f * dni) * nn[1], -nl[2] - (2.0f * dni) * nn[2]};
        const float cdi = 1.0f / sqrtf((cd[0] * cd[0] + cd[1] * cd[1]) + cd[2] * cd[2]);
        const float vdr = ((cd[0] * cdi) * refl[0] + (cd[1] * cdi) * refl[1]) + (cd[2] * cdi) * refl[2];
        const float base = vdr > 0.0f ? vdr : 0.0f;
        spec = pow300(base);
        spec = spec < 0.0f ? 0.0f : (spec > 1.0f ? 1.0f : spec);
    }
    for (int c = 0; c < 3; ++c) {
        float l = 0.33f * diffuse[c];
        l = l + ((0.73f * diffuse[c]) * 0.66f) * intensity;
        l = l + 1.0f * spec;
        Lo[c] = l;
    }
}

// Renders one view.  rgba: H*W*4 bytes; depth (optional): H*W floats = view-space w of the visible fragment, 0 where
// nothing was drawn (V4R's own depth output: uber.vert:103-105, uber.frag:180-182, cleared to 0).
inline void renderView(const Mat4 &view, const std::vector<Instance> &instances, int W, int H, uint8_t *rgba, float *depth) {
    const Projection proj(W, H);
    std::vector<SetupTri> tris;
    for (const Instance &inst : instances) {
        const Mat4 mv = mul(view, inst.model);
        float nm[3][3];
        normalMatrix(mv, nm);
        const MeshRef mesh = meshRef(inst.mesh);
        std::vector<ClipVert> verts(size_t(mesh.nv));
        for (int v = 0; v < mesh.nv; ++v) {
            const Vec3 p{bitsToFloat(mesh.vtx[v][0]), bitsToFloat(mesh.vtx[v][1]), bitsToFloat(mesh.vtx[v][2])};
            const Vec3 nrm{bitsToFloat(mesh.vtx[v][3]), bitsToFloat(mesh.vtx[v][4]), bitsToFloat(mesh.vtx[v][5])};
            const Vec3 cam = transformPoint(mv, p);
            ClipVert &cv = verts[size_t(v)];
            cv.px = cam.x; cv.py = cam.y; cv.pz = cam.z;
            cv.cx = cam.x * proj.p00;
            cv.cy = cam.y * proj.p11;
            cv.cz = cam.z * proj.p22 + proj.p32;
            cv.cw = -cam.z;
            cv.nx = nm[0][0] * nrm.x + nm[1][0] * nrm.y + nm[2][0] * nrm.z;
            cv.ny = nm[0][1] * nrm.x + nm[1][1] * nrm.y + nm[2][1] * nrm.z;
            cv.nz = nm[0][2] * nrm.x + nm[1][2] * nrm.y + nm[2][2] * nrm.z;
        }
        for (int i = 0; i + 2 < mesh.ni; i += 3) {
            const ClipVert tri[3] = {verts[mesh.idx[i]], verts[mesh.idx[i + 1]], verts[mesh.idx[i + 2]]};
            clipAndSetup(tri, W, H, inst.color, tris);
        }
    }

    std::vector<float> zbuf(size_t(W) * H, 1.0f);
    std::vector<int> winner(size_t(W) * H, -1);
    std::vector<float> bary(size_t(W) * H * 3, 0.0f);
    for (int ti = 0; ti < int(tris.size()); ++ti) {
        const SetupTri &t = tris[size_t(ti)];
        int32_t minx = std::min(t.x[0], std::min(t.x[1], t.x[2])), maxx = std::max(t.x[0], std::max(t.x[1], t.x[2]));
        int32_t miny = std::min(t.y[0], std::min(t.y[1], t.y[2])), maxy = std::max(t.y[0], std::max(t.y[1], t.y[2]));
        // pixel px has its centre at px*256+128
        int px0 = std::max(0, (minx - 128 + 255) >> 8), px1 = std::min(W - 1, (maxx - 128) >> 8);
        int py0 = std::max(0, (miny - 128 + 255) >> 8), py1 = std::min(H - 1, (maxy - 128) >> 8);
        // edges: e0 = v1->v2 (weight of v0), e1 = v2->v0 (v1), e2 = v0->v1 (v2)
        bool topleft[3];
        int64_t A[3], B[3], C[3];
        for (int e = 0; e < 3; ++e) {
            const int a = (e + 1) % 3, b = (e + 2) % 3;
            const int64_t dx = int64_t(t.x[b]) - t.x[a], dy = int64_t(t.y[b]) - t.y[a];
            // F(p) = -((bx-ax)(py-ay) - (by-ay)(px-ax)) = dy*px - dx*py + (dx*ay - dy*ax)
            A[e] = dy; B[e] = -dx; C[e] = dx * t.y[a] - dy * t.x[a];
            topleft[e] = (dy == 0 && dx < 0) || dy > 0;
        }
        const float invArea = 1.0f / float(t.area);
        for (int py = py0; py <= py1; ++py)
            for (int px = px0; px <= px1; ++px) {
                const int64_t sx = int64_t(px) * 256 + 128, sy = int64_t(py) * 256 + 128;
                int64_t F[3];
                bool inside = true;
                for (int e = 0; e < 3; ++e) {
                    F[e] = A[e] * sx + B[e] * sy + C[e];
                    if (F[e] < 0 || (F[e] == 0 && !topleft[e])) { inside = false; break; }
                }
                if (!inside) continue;
                const float l0 = float(F[0]) * invArea, l1 = float(F[1]) * invArea, l2 = float(F[2]) * invArea;
                const float z = (l0 * t.z[0] + l1 * t.z[1]) + l2 * t.z[2];
                const size_t pi = size_t(py) * W + px;
                if (z <= zbuf[pi]) {
                    zbuf[pi] = z;
                    winner[pi] = ti;
                    bary[pi * 3 + 0] = l0; bary[pi * 3 + 1] = l1; bary[pi * 3 + 2] = l2;
                }
            }
    }

    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t pi = size_t(py) * W + px;
            uint8_t *o = rgba + pi * 4;
            if (winner[pi] < 0) {
                o[0] = 0; o[1] = 0; o[2] = 0; o[3] = 255;
                if (depth) depth[pi] = 0.0f;
                continue;
            }
            const SetupTri &t = tris[size_t(winner[pi])];
            const float k0 = bary[pi * 3 + 0] * t.rw[0], k1 = bary[pi * 3 + 1] * t.rw[1], k2 = bary[pi * 3 + 2] * t.rw[2];
            const float s = (k0 + k1) + k2;
            const float r = 1.0f / s;
            const float q0 = k0 * r, q1 = k1 * r, q2 = k2 * r;
            float P[3], N[3];
            for (int c = 0; c < 3; ++c) {
                P[c] = (q0 * t.p[0][c] + q1 * t.p[1][c]) + q2 * t.p[2][c];
                N[c] = (q0 * t.n[0][c] + q1 * t.n[1][c]) + q2 * t.n[2][c];
            }
            if (depth) depth[pi] = r;  // interpolated gl_Position.w == 1 / (sum lambda_i / w_i)
            const uint32_t rgb = uint32_t(allColors[t.color]);
            const float diffuse[3] = {float((rgb >> 16) & 255) / 255.0f, float((rgb >> 8) & 255) / 255.0f, float(rgb & 255) / 255.0f};
            float Lo[3];
            shadeFragment(P, N, diffuse, Lo);
            for (int c = 0; c < 3; ++c) o[c] = toUnorm8(Lo[c]);
            o[3] = 255;
        }
}

}  // namespace orc
