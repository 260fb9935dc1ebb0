// K4: batched first-person rasteriser, two kernels per view chunk.
//
//   geomKernel   one thread per (view, instance face | mesh triangle): model-view transform, near/far clip, projection,
//                8-bit sub-pixel snap, back-face cull, edge/plane set-up -> per-view triangle list in global scratch
//                (a few MB per chunk: stays L2 resident, never meant to reach HBM)
//   tileKernel   one warp per 32x4-pixel tile of a view: lanes scan the view's triangle boxes 32 at a time (ballot), the
//                surviving triangles are evaluated with exact integer edge functions (top-left rule), perspective-correct
//                varyings, nearest fragment kept in registers (depth LESS_OR_EQUAL, later draw wins ties), deferred
//                Phong-clone shading, packed RGBA8 written with one 128-bit store per lane (8 lanes = one 128-byte line)
//
// Replaces (file:line under /root/reference):
//   V4R CommandStreamState::render        src/3rdparty/v4r/src/vulkan_state.inl:10-160
//   vertex / fragment shaders             src/3rdparty/v4r/src/pipelines/shaders/uber.vert:53-110, uber.frag:112-141
//   projection                            src/3rdparty/v4r/src/v4r.cpp:35-45
//   raster + depth state                  src/3rdparty/v4r/src/vulkan_state.cpp:588-606
//   image -> linear buffer copy           src/3rdparty/v4r/src/vulkan_state.cpp:909-957 (obs layout uint8[N][H][W][4])
//   instance lists, draw order            src/libs/v4r_rendering/src/v4r_env_renderer.cpp:267-279
#pragma once
#include "dev_math.cuh"
#include "mesh_tables.inc"
#include "mv_types.h"

namespace mvr {
using namespace dm;

struct __align__(16) TriCover {  // what the coverage / depth loop reads (broadcast loads)
    long long C[3];     // edge constant terms, top-left bias already applied
    int32_t A[3], B[3];
    float z[3];
    float invArea;
    uint32_t key;       // draw order + 1, < 2^18 (later wins depth ties: LESS_OR_EQUAL)
    int32_t flags;      // bits 0..2: edge e is top-left (no bias was applied); bit 3: every edge function fits int32 in the viewport
    uint32_t bx, by;    // pixel box, inclusive: x0 | x1 << 16, y0 | y1 << 16
};
struct __align__(16) TriShade {  // what deferred shading reads for the winning fragment
    float rw[3];
    float p[9];
    float n[9];
    int32_t color;
    int32_t pad[2];
};
static_assert(sizeof(TriCover) == 80 && sizeof(TriShade) == 96, "triangle record layout");

struct RasterParams {
    const MvInstance *instances; // [E][instStride] drawables in draw order (boxes first)
    const int32_t *instCounts;   // [E][8] {boxes, total, capsules, spheres, cones, cylinders, -, -}
    const float *views;          // [E*A][16]
    int instStride;
    uint8_t *obs;                // [N][H][W][4]
    float *depth;                // [N][H][W] or nullptr
    int32_t *faults;             // [E] (ORed)
    // triangle scratch for views [viewBase, viewBase + chunkViews)
    TriCover *cover;             // [chunkViews][triCap]
    TriShade *shade;             // [chunkViews][triCap]
    // binning: a triangle whose box touches at most kWideTiles tiles is appended to each of those tiles' bins; wider ones
    // go to the view's wide list (box + index), which every tile of the view checks
    int32_t *binCounts;          // [chunkViews][tiles]; every tile warp zeroes its own counter after reading it
    uint16_t *binList;           // [chunkViews][tiles][binCap] triangle indices
    int32_t *wideCounts;         // [N], zeroed before geomKernel (the step kernel does it)
    int4 *wideList;              // [chunkViews][kWideCap] {bx, by, index, 0}
    int binCap;
    int32_t *triCounts;          // [N], zeroed before geomKernel (the step kernel does it)
    int32_t *tileCounter;        // dynamic tile queue of the persistent tile warps (reset by geomKernel)
    const uint32_t *ready;       // [E] step-kernel completion stamps (nullptr: plain stream order)
    uint32_t readyStamp;         // value ready[env] holds once this step's state, instances and views of env are written
    uint32_t *sliceDone;         // optional [slices] cumulative count of finished tiles per slice of sliceViews views: lets a copy
    int sliceViews;              //   stream (cuStreamWaitValue32) download finished slices while later views are still rasterised
    uint32_t *tileProf;          // optional [views][tiles][4] {cycles, triangles overlapping, small-path lanes, big-path triangles} (debug)
    int fastShading;             // 1: approximate rsqrt / fused multiply-add in the fragment stage (+-1 LSB), 0: bit-exact
    int viewBase, chunkViews;
    int N, A, W, H;
    int triCap;
    float p00, p11, p22, p32;
};

struct ClipVert { float cx, cy, cz, cw, px, py, pz, nx, ny, nz; };

__device__ __forceinline__ ClipVert lerpVert(const ClipVert &a, const ClipVert &b, float t) {
    ClipVert o;
    o.cx = a.cx + t * (b.cx - a.cx); o.cy = a.cy + t * (b.cy - a.cy); o.cz = a.cz + t * (b.cz - a.cz); o.cw = a.cw + t * (b.cw - a.cw);
    o.px = a.px + t * (b.px - a.px); o.py = a.py + t * (b.py - a.py); o.pz = a.pz + t * (b.pz - a.pz);
    o.nx = a.nx + t * (b.nx - a.nx); o.ny = a.ny + t * (b.ny - a.ny); o.nz = a.nz + t * (b.nz - a.nz);
    return o;
}
__device__ __forceinline__ int32_t snapSub(float v) { return int32_t(floorf(v * 256.0f + 0.5f)); }

#ifndef MV_TILE_RUN
#define MV_TILE_RUN 2
#endif
constexpr int kTileRun = MV_TILE_RUN;  // consecutive tiles per queue claim of the tile kernel
constexpr int kWideTiles = 8;   // more tiles than this: the triangle goes to the view's wide list
constexpr int kWideCap = 512;

struct SetupCtx {
    TriCover *cover;
    TriShade *shade;
    int32_t *binCounts;   // this view's [tiles]
    uint16_t *binList;    // this view's [tiles][binCap]
    int32_t *wideCount;
    int4 *wideList;
    int binCap;
    int32_t *nTris;
    int32_t *fault;
    int triCap;
    int W, H;
};

// one projected triangle: cull, box, edge / plane set-up, append to the view's list
__device__ __forceinline__ void emitTri(const SetupCtx &cx, const ClipVert &va, const ClipVert &vb, const ClipVert &vc, const int32_t sxs[3], const int32_t sys[3],
                                        const float szs[3], const float rws[3], int color, uint32_t key) {
    const long long area2 = (long long)(sxs[1] - sxs[0]) * (long long)(sys[2] - sys[0]) - (long long)(sys[1] - sys[0]) * (long long)(sxs[2] - sxs[0]);
    if (area2 >= 0) return;  // back-facing (visually clockwise with y down) or degenerate
    const int32_t minx = min(sxs[0], min(sxs[1], sxs[2])), maxx = max(sxs[0], max(sxs[1], sxs[2]));
    const int32_t miny = min(sys[0], min(sys[1], sys[2])), maxy = max(sys[0], max(sys[1], sys[2]));
    const int px0 = max(0, (minx - 128 + 255) >> 8), px1 = min(cx.W - 1, (maxx - 128) >> 8);
    const int py0 = max(0, (miny - 128 + 255) >> 8), py1 = min(cx.H - 1, (maxy - 128) >> 8);
    if (px0 > px1 || py0 > py1) return;  // covers no pixel centre of the viewport
    const int slot = atomicAdd(cx.nTris, 1);
    if (slot >= cx.triCap) { atomicOr(cx.fault, MV_FAULT_TRI_OVERFLOW); return; }
    TriCover c;
    TriShade s;
    const ClipVert *vs[3] = {&va, &vb, &vc};
    int tl = 0;
    long long worst = 0;
    const long long wsub = (long long)cx.W * 256, hsub = (long long)cx.H * 256;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int a = (e + 1) % 3, b = (e + 2) % 3;
        const long long dx = (long long)sxs[b] - sxs[a], dy = (long long)sys[b] - sys[a];
        const bool topleft = (dy == 0 && dx < 0) || dy > 0;
        c.A[e] = int32_t(dy);
        c.B[e] = int32_t(-dx);
        c.C[e] = dx * sys[a] - dy * sxs[a] - (topleft ? 0 : 1);
        tl |= topleft ? (1 << e) : 0;
        const long long bound = llabs(dy) * wsub + llabs(dx) * hsub + llabs(c.C[e]);
        worst = bound > worst ? bound : worst;
        c.z[e] = szs[e]; s.rw[e] = rws[e];
        s.p[e * 3 + 0] = vs[e]->px; s.p[e * 3 + 1] = vs[e]->py; s.p[e * 3 + 2] = vs[e]->pz;
        s.n[e * 3 + 0] = vs[e]->nx; s.n[e * 3 + 1] = vs[e]->ny; s.n[e * 3 + 2] = vs[e]->nz;
    }
    c.invArea = 1.0f / float(-area2);
    c.key = key;
    c.flags = tl | (worst < (1ll << 30) ? 8 : 0);  // bit 3: every edge function fits int32 anywhere in the viewport
    c.bx = uint32_t(px0) | (uint32_t(px1) << 16);
    c.by = uint32_t(py0) | (uint32_t(py1) << 16);
    s.color = color; s.pad[0] = 0; s.pad[1] = 0;
    cx.cover[slot] = c;
    cx.shade[slot] = s;
    // bin it
    const int tilesX = cx.W >> 5;
    const int tx0 = px0 >> 5, tx1 = px1 >> 5, ty0 = py0 >> 2, ty1 = py1 >> 2;
    if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > kWideTiles) {
        const int w = atomicAdd(cx.wideCount, 1);
        if (w < kWideCap) cx.wideList[w] = make_int4(int(c.bx), int(c.by), slot, 0);
        else atomicOr(cx.fault, MV_FAULT_TRI_OVERFLOW);
    } else {
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                const int t = ty * tilesX + tx;
                const int b = atomicAdd(cx.binCounts + t, 1);
                if (b < cx.binCap) cx.binList[size_t(t) * cx.binCap + b] = uint16_t(slot);
                else atomicOr(cx.fault, MV_FAULT_TRI_OVERFLOW);
            }
    }
}

// clip against z >= 0 and z <= w, project, snap, cull, emit
__device__ void clipAndSetup(const SetupCtx &cx, const ClipVert &v0, const ClipVert &v1, const ClipVert &v2, int color, uint32_t keyBase) {
    const float hw = float(cx.W) * 0.5f, hh = float(cx.H) * 0.5f;
    const bool allIn = v0.cz >= 0.0f && v1.cz >= 0.0f && v2.cz >= 0.0f && (v0.cw - v0.cz) >= 0.0f && (v1.cw - v1.cz) >= 0.0f && (v2.cw - v2.cz) >= 0.0f;
    if (allIn) {  // the common case stays in registers
        const ClipVert *vs[3] = {&v0, &v1, &v2};
        int32_t sxs[3], sys[3];
        float szs[3], rws[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float r = 1.0f / vs[i]->cw;
            rws[i] = r;
            sxs[i] = snapSub((vs[i]->cx * r) * hw + hw);
            sys[i] = snapSub((vs[i]->cy * r) * hh + hh);
            szs[i] = vs[i]->cz * r;
        }
        emitTri(cx, v0, v1, v2, sxs, sys, szs, rws, color, keyBase);
        return;
    }
    ClipVert poly[6], tmp[6];
    int n = 3;
    poly[0] = v0; poly[1] = v1; poly[2] = v2;
    for (int plane = 0; plane < 2; ++plane) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const ClipVert &a = poly[i];
            const ClipVert &b = poly[(i + 1) % n];
            const float da = plane == 0 ? a.cz : a.cw - a.cz;
            const float db = plane == 0 ? b.cz : b.cw - b.cz;
            const bool ina = da >= 0.0f, inb = db >= 0.0f;
            if (ina) tmp[m++] = a;
            if (ina != inb) {
                if (ina) tmp[m++] = lerpVert(a, b, da / (da - db));
                else tmp[m++] = lerpVert(b, a, db / (db - da));
            }
        }
        n = m;
        for (int i = 0; i < n; ++i) poly[i] = tmp[i];
        if (n < 3) return;
    }
    int32_t sx[6], sy[6];
    float sz[6], rw[6];
    for (int i = 0; i < n; ++i) {
        const float r = 1.0f / poly[i].cw;
        rw[i] = r;
        sx[i] = snapSub((poly[i].cx * r) * hw + hw);
        sy[i] = snapSub((poly[i].cy * r) * hh + hh);
        sz[i] = poly[i].cz * r;
    }
    for (int k = 1; k + 1 < n; ++k) {
        const int32_t sxs[3] = {sx[0], sx[k], sx[k + 1]}, sys[3] = {sy[0], sy[k], sy[k + 1]};
        const float szs[3] = {sz[0], sz[k], sz[k + 1]}, rws[3] = {rw[0], rw[k], rw[k + 1]};
        emitTri(cx, poly[0], poly[k], poly[k + 1], sxs, sys, szs, rws, color, keyBase);
    }
}

__device__ __forceinline__ ClipVert makeVert(const M4 &mv, const float nm[9], V3 p, V3 nrm, float p00, float p11, float p22, float p32) {
    const V3 cam = transformPoint(mv, p);
    ClipVert cv;
    cv.px = cam.x; cv.py = cam.y; cv.pz = cam.z;
    cv.cx = cam.x * p00;
    cv.cy = cam.y * p11;
    cv.cz = cam.z * p22 + p32;
    cv.cw = -cam.z;
    cv.nx = nm[0] * nrm.x + nm[3] * nrm.y + nm[6] * nrm.z;
    cv.ny = nm[1] * nrm.x + nm[4] * nrm.y + nm[7] * nrm.z;
    cv.nz = nm[2] * nrm.x + nm[5] * nrm.y + nm[8] * nrm.z;
    return cv;
}

// ---------------------------------------------------------------------------------------------------- geometry kernel
// grid = (ceil(maxItems / blockDim.x), chunkViews); blockIdx.y selects the view
#ifndef MV_GEOM_BLOCKS
#define MV_GEOM_BLOCKS 4
#endif
// Conservative instance-level frustum test (option "cull", off by default): the bounding sphere of the instance's mesh in view space
// against the near plane and the four side planes.  Unit meshes span |x|,|z| <= 1 and |y| <= 1 (capsule: 2), so a vertex lies within
// bx*|col0| + by*|col1| + bz*|col2| of the instance origin.  An instance that fails contributes no fragment (clipAndSetup would have
// clipped or scissored every one of its triangles), so frames do not change.
__device__ __forceinline__ bool instanceMayBeVisible(const M4 &mv, float by, float p00, float p11) {
    const float l0 = sqrtf(mv.c[0] * mv.c[0] + mv.c[1] * mv.c[1] + mv.c[2] * mv.c[2]);
    const float l1 = sqrtf(mv.c[4] * mv.c[4] + mv.c[5] * mv.c[5] + mv.c[6] * mv.c[6]);
    const float l2 = sqrtf(mv.c[8] * mv.c[8] + mv.c[9] * mv.c[9] + mv.c[10] * mv.c[10]);
    const float r = (l0 + by * l1 + l2) * 1.001f + 1e-3f;
    const float x = mv.c[12], y = mv.c[13], z = mv.c[14];
    if (-z - 0.01f < -r) return false;  // wholly in front of the near plane (camera looks down -z)
    // side planes x_clip = +-w_clip, y_clip = +-w_clip with x_clip = p00 * x, y_clip = p11 * y, w_clip = -z: inward unit normals
    const float ix = rsqrtf(p00 * p00 + 1.0f), iy = rsqrtf(p11 * p11 + 1.0f);
    const float ax = fabsf(p00), ay = fabsf(p11);
    if ((-ax * x - z) * ix < -r || (ax * x - z) * ix < -r) return false;
    if ((-ay * y - z) * iy < -r || (ay * y - z) * iy < -r) return false;
    return true;
}

template <bool CULL>
__global__ void __launch_bounds__(128, MV_GEOM_BLOCKS) geomKernel(RasterParams P) {
    const int vslot = blockIdx.y;
    const int view = P.viewBase + vslot;
    if (view >= P.N) return;
    const int env = view / P.A;
    int item = blockIdx.x * blockDim.x + threadIdx.x;
    // CULL: per-block tables of the (at most 23) instances this block's 128 items belong to, filled by one thread per instance
    __shared__ float s_mv[CULL ? 24 : 1][16];
    __shared__ float s_nm[CULL ? 24 : 1][9];
    __shared__ int s_color[CULL ? 24 : 1];
    int cullFirst = 0;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *P.tileCounter = 0;  // for the tile kernel that follows in-stream
    // Launched with programmatic stream serialisation this grid starts while the step kernel is still running: each block
    // waits for its own env's completion stamp (release/acquire through L2) instead of for the whole step grid, so the
    // geometry of finished envs overlaps the step kernel's long-tail envs.  Everything the step kernel produced is then
    // read with L2-coherent loads (ld.global.cg), never through the non-coherent path.
    if (P.ready) {
        if (threadIdx.x == 0) {
            const uint32_t *flag = P.ready + env;
            uint32_t v;
            while (true) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (v == P.readyStamp) break;
                __nanosleep(200);
            }
        }
        __syncthreads();
    }
    const MvInstance *inst = P.instances + size_t(env) * P.instStride;
    // work items: 6 faces per box, then one per triangle of the other meshes (instances are sorted by mesh type)
    int cnt[6];
    {
        const int4 c03 = __ldcg(reinterpret_cast<const int4 *>(P.instCounts + env * 8));
        const int2 c45 = __ldcg(reinterpret_cast<const int2 *>(P.instCounts + env * 8 + 4));
        cnt[0] = c03.x; cnt[1] = c03.y; cnt[2] = c03.z; cnt[3] = c03.w; cnt[4] = c45.x; cnt[5] = c45.y;
    }
    const int nBoxInst = cnt[0];
    const int nBoxItems = nBoxInst * 6;
    const int capItems = cnt[2] * MV_CAPSULE_TRIS, sphItems = cnt[3] * MV_SPHERE_TRIS, coneItems = cnt[4] * MV_CONE_TRIS, cylItems = cnt[5] * MV_CYLINDER_TRIS;
    if (!CULL) {
        if (item >= nBoxItems + capItems + sphItems + coneItems + cylItems) return;
    }

    SetupCtx cx;
    cx.cover = P.cover + size_t(vslot) * P.triCap;
    cx.shade = P.shade + size_t(vslot) * P.triCap;
    {
        const int nTilesV = (P.W >> 5) * (P.H >> 2);
        cx.binCounts = P.binCounts + size_t(vslot) * nTilesV;
        cx.binList = P.binList + size_t(vslot) * nTilesV * P.binCap;
        cx.wideCount = P.wideCounts + view;
        cx.wideList = P.wideList + size_t(vslot) * kWideCap;
        cx.binCap = P.binCap;
    }
    cx.nTris = P.triCounts + view;
    cx.fault = P.faults + env;
    cx.triCap = P.triCap; cx.W = P.W; cx.H = P.H;

    M4 viewM;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 col = __ldcg(reinterpret_cast<const float4 *>(P.views + size_t(view) * 16) + i);
        viewM.c[i * 4 + 0] = col.x; viewM.c[i * 4 + 1] = col.y; viewM.c[i * 4 + 2] = col.z; viewM.c[i * 4 + 3] = col.w;
    }

    if (CULL) {
        // block-level cull + compaction: the first thread of every instance inside this block's 128 items tests the instance, the
        // survivors' items are packed to the front, and whole warps beyond them retire before any vertex work
        __shared__ uint8_t s_vis[128];
        __shared__ uint8_t s_items[128];
        __shared__ int s_warpCount[4];
        const int total = nBoxItems + capItems + sphItems + coneItems + cylItems;
        if (int(blockIdx.x * blockDim.x) >= total) return;  // block-uniform
        auto instanceOf = [&](int it, int &sub, float &by) {
            by = 1.0f;
            if (it < nBoxItems) { sub = it % 6; return it / 6; }
            int rest = it - nBoxItems;
            if (rest < capItems) { sub = rest % MV_CAPSULE_TRIS; by = 2.0f; return nBoxInst + rest / MV_CAPSULE_TRIS; }
            if ((rest -= capItems) < sphItems) { sub = rest % MV_SPHERE_TRIS; return nBoxInst + cnt[2] + rest / MV_SPHERE_TRIS; }
            if ((rest -= sphItems) < coneItems) { sub = rest % MV_CONE_TRIS; return nBoxInst + cnt[2] + cnt[3] + rest / MV_CONE_TRIS; }
            rest -= coneItems;
            sub = rest % MV_CYLINDER_TRIS;
            return nBoxInst + cnt[2] + cnt[3] + cnt[4] + rest / MV_CYLINDER_TRIS;
        };
        const bool valid = item < total;
        int sub0 = 0, sub = 0;
        float by0 = 1.0f, by = 1.0f;
        const int ii0 = instanceOf(int(blockIdx.x * blockDim.x), sub0, by0);
        const int ii = valid ? instanceOf(item, sub, by) : ii0;
        if (valid && (sub == 0 || threadIdx.x == 0)) {
            M4 model;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 col = __ldcg(reinterpret_cast<const float4 *>(inst[ii].model) + i);
                model.c[i * 4 + 0] = col.x; model.c[i * 4 + 1] = col.y; model.c[i * 4 + 2] = col.z; model.c[i * 4 + 3] = col.w;
            }
            const M4 mvL = mul4(viewM, model);
            const bool vis = instanceMayBeVisible(mvL, by, P.p00, P.p11);
            s_vis[ii - ii0] = vis ? 1 : 0;
            if (vis) {  // the same products every item of this instance would otherwise repeat
                float nmL[9];
                normalMatrix(mvL, nmL);
#pragma unroll
                for (int q = 0; q < 16; ++q) s_mv[ii - ii0][q] = mvL.c[q];
#pragma unroll
                for (int q = 0; q < 9; ++q) s_nm[ii - ii0][q] = nmL[q];
                s_color[ii - ii0] = __ldcg(&inst[ii].color);
            }
        }
        __syncthreads();
        bool alive = valid && s_vis[ii - ii0] != 0;
        if (alive && item < nBoxItems) {
            // a box face whose plane clearly faces away from the eye (at the view-space origin) only yields triangles the
            // rasteriser's winding test drops: outward normal n, face centre c = origin + n (unit cube), cull when n_view . c_view > 0
            const float *m = s_mv[ii - ii0], *nmS = s_nm[ii - ii0];
            const float *fn = c_boxVerts[sub * 4] + 3;
            const float cxv = m[12] + (fn[0] * m[0] + fn[1] * m[4] + fn[2] * m[8]);
            const float cyv = m[13] + (fn[0] * m[1] + fn[1] * m[5] + fn[2] * m[9]);
            const float czv = m[14] + (fn[0] * m[2] + fn[1] * m[6] + fn[2] * m[10]);
            const float nxv = nmS[0] * fn[0] + nmS[3] * fn[1] + nmS[6] * fn[2];
            const float nyv = nmS[1] * fn[0] + nmS[4] * fn[1] + nmS[7] * fn[2];
            const float nzv = nmS[2] * fn[0] + nmS[5] * fn[1] + nmS[8] * fn[2];
            const float d = nxv * cxv + nyv * cyv + nzv * czv;
            if (d > 1e-3f * sqrtf((nxv * nxv + nyv * nyv + nzv * nzv) * (cxv * cxv + cyv * cyv + czv * czv))) alive = false;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, alive);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) s_warpCount[warp] = __popc(bal);
        __syncthreads();
        int base = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) { if (w < warp) base += s_warpCount[w]; tot += s_warpCount[w]; }
        if (alive) s_items[base + __popc(bal & ((1u << lane) - 1u))] = uint8_t(threadIdx.x);
        __syncthreads();
        if (int(threadIdx.x) >= tot) return;
        item = int(blockIdx.x * blockDim.x) + int(s_items[threadIdx.x]);
        cullFirst = ii0;
    }

    if (item < nBoxItems) {
        const int ii = item / 6, face = item % 6;
        M4 mv;
        float nm[9];
        int color;
        if (CULL) {
#pragma unroll
            for (int q = 0; q < 16; ++q) mv.c[q] = s_mv[ii - cullFirst][q];
#pragma unroll
            for (int q = 0; q < 9; ++q) nm[q] = s_nm[ii - cullFirst][q];
            color = s_color[ii - cullFirst];
        } else {
            M4 model;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 col = __ldcg(reinterpret_cast<const float4 *>(inst[ii].model) + i);
                model.c[i * 4 + 0] = col.x; model.c[i * 4 + 1] = col.y; model.c[i * 4 + 2] = col.z; model.c[i * 4 + 3] = col.w;
            }
            color = __ldcg(&inst[ii].color);
            mv = mul4(viewM, model);
            normalMatrix(mv, nm);
        }
        ClipVert cvt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float *vp = c_boxVerts[face * 4 + k];
            cvt[k] = makeVert(mv, nm, v3(vp[0], vp[1], vp[2]), v3(vp[3], vp[4], vp[5]), P.p00, P.p11, P.p22, P.p32);
        }
        // draw-order key: instance, then triangle within the mesh (the fan pieces of one clipped triangle share a key: they are
        // coplanar and disjoint, so they never tie on a pixel).  18 bits + 13 bits of list index fill the fragment's low word.
        const uint32_t keyBase = uint32_t(ii) * 128u + uint32_t(face) * 2u + 1u;
        clipAndSetup(cx, cvt[0], cvt[1], cvt[2], color, keyBase);       // cube indices f*4+{0,1,2}
        clipAndSetup(cx, cvt[0], cvt[2], cvt[3], color, keyBase + 1u);  // cube indices f*4+{0,2,3}
    } else {
        int rest = item - nBoxItems, ii = nBoxInst, tri, mesh;
        if (rest < capItems) { mesh = 1; ii += rest / MV_CAPSULE_TRIS; tri = rest % MV_CAPSULE_TRIS; }
        else if ((rest -= capItems) < sphItems) { mesh = 2; ii += cnt[2] + rest / MV_SPHERE_TRIS; tri = rest % MV_SPHERE_TRIS; }
        else if ((rest -= sphItems) < coneItems) { mesh = 3; ii += cnt[2] + cnt[3] + rest / MV_CONE_TRIS; tri = rest % MV_CONE_TRIS; }
        else { rest -= coneItems; mesh = 4; ii += cnt[2] + cnt[3] + cnt[4] + rest / MV_CYLINDER_TRIS; tri = rest % MV_CYLINDER_TRIS; }
        M4 mv;
        float nm[9];
        if (CULL) {
#pragma unroll
            for (int q = 0; q < 16; ++q) mv.c[q] = s_mv[ii - cullFirst][q];
#pragma unroll
            for (int q = 0; q < 9; ++q) nm[q] = s_nm[ii - cullFirst][q];
        } else {
            M4 model;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 col = __ldcg(reinterpret_cast<const float4 *>(inst[ii].model) + i);
                model.c[i * 4 + 0] = col.x; model.c[i * 4 + 1] = col.y; model.c[i * 4 + 2] = col.z; model.c[i * 4 + 3] = col.w;
            }
            mv = mul4(viewM, model);
            normalMatrix(mv, nm);
        }
        ClipVert cvt[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float *vp;
            if (mesh == 1) vp = c_capsuleVerts[c_capsuleIdx[tri * 3 + k]];
            else if (mesh == 2) vp = c_sphereVerts[c_sphereIdx[tri * 3 + k]];
            else if (mesh == 3) vp = c_coneVerts[c_coneIdx[tri * 3 + k]];
            else vp = c_cylinderVerts[c_cylinderIdx[tri * 3 + k]];
            cvt[k] = makeVert(mv, nm, v3(vp[0], vp[1], vp[2]), v3(vp[3], vp[4], vp[5]), P.p00, P.p11, P.p22, P.p32);
        }
        const uint32_t keyBase = uint32_t(ii) * 128u + uint32_t(tri) + 1u;
        clipAndSetup(cx, cvt[0], cvt[1], cvt[2], __ldcg(&inst[ii].color), keyBase);
    }
}

// ---------------------------------------------------------------------------------------------------- shading
__device__ __forceinline__ float pow300(float x) {
    const float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8, x32 = x16 * x16, x64 = x32 * x32, x128 = x64 * x64, x256 = x128 * x128;
    return ((x256 * x32) * x8) * x4;
}
__device__ __forceinline__ uint32_t toUnorm8(float c) {
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    return uint32_t(floorf(c * 255.0f + 0.5f));
}

__constant__ float c_palette[22][3];

template <bool FAST> __device__ __forceinline__ float invLen3(float x, float y, float z) {
    if (FAST) return rsqrtf(__fmaf_rn(z, z, __fmaf_rn(y, y, x * x)));
    return 1.0f / sqrtf((x * x + y * y) + z * z);
}
template <bool FAST> __device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    if (FAST) return __fmaf_rn(az, bz, __fmaf_rn(ay, by, ax * bx));
    return (ax * bx + ay * by) + az * bz;
}

// uber.frag:112-141.  FAST keeps the structure but uses rsqrt.approx + FMA: colours move by at most 1 LSB (the tolerance the
// north star grants for RGB); the exact variant reproduces the oracle byte for byte.  The depth output is exact in both.
struct ShadeRec { float4 a0, a1, a2, a3, a4, a5; };
__device__ __forceinline__ ShadeRec loadShade(const TriShade *tp) {
    const float4 *q = reinterpret_cast<const float4 *>(tp);  // 96-byte record as six 128-bit read-only loads
    ShadeRec r;
    r.a0 = __ldg(q + 0); r.a1 = __ldg(q + 1); r.a2 = __ldg(q + 2); r.a3 = __ldg(q + 3); r.a4 = __ldg(q + 4); r.a5 = __ldg(q + 5);
    return r;
}
template <bool FAST> __device__ __forceinline__ uint32_t shadePixel(const ShadeRec &rec, float l0, float l1, float l2, float &wOut) {
    const float4 a0 = rec.a0, a1 = rec.a1, a2 = rec.a2, a3 = rec.a3, a4 = rec.a4, a5 = rec.a5;
    const float rw0 = a0.x, rw1 = a0.y, rw2 = a0.z;
    const float p[9] = {a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
    const float n[9] = {a3.x, a3.y, a3.z, a3.w, a4.x, a4.y, a4.z, a4.w, a5.x};
    const int color = __float_as_int(a5.y);
    const float k0 = l0 * rw0, k1 = l1 * rw1, k2 = l2 * rw2;
    const float s = (k0 + k1) + k2;
    const float r = 1.0f / s;
    const float q0 = k0 * r, q1 = k1 * r, q2 = k2 * r;
    float Pc[3], N[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        Pc[c] = dot3<FAST>(q0, q1, q2, p[c], p[3 + c], p[6 + c]);
        N[c] = dot3<FAST>(q0, q1, q2, n[c], n[3 + c], n[6 + c]);
    }
    wOut = r;
    const float cd0 = -Pc[0], cd1 = -Pc[1], cd2 = -Pc[2];
    const float ld0 = 0.0f + cd0, ld1 = 4.0f + cd1, ld2 = 2.0f + cd2;
    const float ldi = invLen3<FAST>(ld0, ld1, ld2);
    const float nl0 = ld0 * ldi, nl1 = ld1 * ldi, nl2 = ld2 * ldi;
    const float nni = invLen3<FAST>(N[0], N[1], N[2]);
    const float nn0 = N[0] * nni, nn1 = N[1] * nni, nn2 = N[2] * nni;
    const float ndl = dot3<FAST>(nn0, nn1, nn2, nl0, nl1, nl2);
    const float intensity = ndl > 0.0f ? ndl : 0.0f;
    float spec = 0.0f;
    if (intensity > 0.001f) {
        const float dni = -ndl;
        const float r0 = -nl0 - (2.0f * dni) * nn0, r1 = -nl1 - (2.0f * dni) * nn1, r2 = -nl2 - (2.0f * dni) * nn2;
        const float cdi = invLen3<FAST>(cd0, cd1, cd2);
        const float vdr = dot3<FAST>(cd0 * cdi, cd1 * cdi, cd2 * cdi, r0, r1, r2);
        const float base = vdr > 0.0f ? vdr : 0.0f;
        spec = pow300(base);
        spec = spec < 0.0f ? 0.0f : (spec > 1.0f ? 1.0f : spec);
    }
    uint32_t out = 0xff000000u;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float diffuse = c_palette[color][c];
        float Lo;
        if (FAST) {
            Lo = __fmaf_rn((0.73f * diffuse) * 0.66f, intensity, 0.33f * diffuse) + spec;
            out |= __float2uint_rn(__saturatef(Lo) * 255.0f) << (8 * c);
        } else {
            Lo = 0.33f * diffuse;
            Lo = Lo + ((0.73f * diffuse) * 0.66f) * intensity;
            Lo = Lo + 1.0f * spec;
            out |= toUnorm8(Lo) << (8 * c);
        }
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------- tile raster kernel
// Persistent warps pull (view, 32x4 tile) work items from a global counter.  Inside a tile the view's triangle list is
// scanned 32 boxes at a time; triangles that touch the tile take one of two paths:
//   * small (at most kSmallArea pixels of the tile): ONE LANE PER TRIANGLE walks its few pixels and publishes fragments
//     with a packed 64-bit shared-memory atomicMax  ->  dense clusters of tiny triangles cost ~1/32 of the serial path
//   * large: the whole warp evaluates it, lane = 4 horizontally adjacent pixels, best fragment kept in registers
// A fragment is the 64-bit key (~depth bits << 32) | (draw order << 13) | list index: max == nearest depth, and on equal
// depth the later draw (LESS_OR_EQUAL).  Barycentrics are recomputed for the single winner at shading time.
constexpr int kSmallArea = 24;

struct EdgeEval {  // one triangle's edge functions at a pixel centre
    int A0, A1, A2, B0, B1, B2, u0, u1, u2, small;
    long long C0, C1, C2;
    float z0, z1, z2, invArea;
    uint32_t key;
};
__device__ __forceinline__ EdgeEval unpackCover(const int4 q0, const int4 q1, const int4 q2, const int4 q3, const int4 q4) {
    EdgeEval e;
    e.C0 = (long long)(((unsigned long long)(unsigned)q0.y << 32) | (unsigned)q0.x);
    e.C1 = (long long)(((unsigned long long)(unsigned)q0.w << 32) | (unsigned)q0.z);
    e.C2 = (long long)(((unsigned long long)(unsigned)q1.y << 32) | (unsigned)q1.x);
    e.A0 = q1.z; e.A1 = q1.w; e.A2 = q2.x; e.B0 = q2.y; e.B1 = q2.z; e.B2 = q2.w;
    e.z0 = __int_as_float(q3.x); e.z1 = __int_as_float(q3.y); e.z2 = __int_as_float(q3.z); e.invArea = __int_as_float(q3.w);
    e.key = uint32_t(q4.x);
    const int fl = q4.y;
    e.u0 = (fl & 1) ? 0 : 1; e.u1 = (fl & 2) ? 0 : 1; e.u2 = (fl & 4) ? 0 : 1;  // undo the top-left bias for the barycentrics
    e.small = (fl >> 3) & 1;
    return e;
}
__device__ __forceinline__ EdgeEval loadCover(const TriCover *c) {  // global memory, read-only path
    const int4 *cq = reinterpret_cast<const int4 *>(c);
    return unpackCover(__ldg(cq + 0), __ldg(cq + 1), __ldg(cq + 2), __ldg(cq + 3), __ldg(cq + 4));
}
__device__ __forceinline__ EdgeEval loadCoverShared(const TriCover *c) {
    const int4 *cq = reinterpret_cast<const int4 *>(c);
    return unpackCover(cq[0], cq[1], cq[2], cq[3], cq[4]);
}
// the fragment's low word is (draw-order key << 13) | triangle-list index: key = instance * 128 + triangle-in-mesh + 1 has to stay below
// 2^19 and the per-view triangle capacity below 2^13 -- the constraint to revisit when a capacity in mv_types.h is raised
static_assert((MV_MAX_INSTANCES + 1) * 128 <= (1 << 19), "draw-order key does not fit beside the 13-bit triangle index");
static_assert(MV_CAPSULE_TRIS <= 128 && MV_SPHERE_TRIS <= 128 && MV_CONE_TRIS <= 128 && MV_CYLINDER_TRIS <= 128, "triangle-in-mesh index needs 7 bits");
__device__ __forceinline__ unsigned long long packFrag(float z, uint32_t key, int idx) {
    const uint32_t b = __float_as_uint(z);
    const uint32_t asc = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);  // monotonic in z over all floats (tiny negative z can come out of the clipper)
    return ((unsigned long long)(~asc) << 32) | (unsigned long long)((key << 13) | uint32_t(idx));
}

#ifndef MV_TILE_BLOCKS
#define MV_TILE_BLOCKS 5
#endif
constexpr int kTileBlocksPerSM = MV_TILE_BLOCKS;  // persistent blocks of 4 warps per SM (bounds the register budget)
template <bool FAST> __global__ void __launch_bounds__(128, kTileBlocksPerSM) tileKernel(RasterParams P) {
    __shared__ unsigned long long s_frag[4][128];
    __shared__ __align__(16) TriCover s_stage[4][32];  // this chunk's triangle records, one per lane (the large path reads them back)
    const int lane = threadIdx.x & 31;
    unsigned long long *frag = s_frag[threadIdx.x >> 5];
    TriCover *stage = s_stage[threadIdx.x >> 5];
    const int tilesX = P.W / 32, nTiles = tilesX * (P.H / 4);
    const int totalTiles = min(P.chunkViews, P.N - P.viewBase) * nTiles;
    // Work queue: one global counter claimed in short runs of consecutive tiles (fewer same-address atomics, and neighbouring
    // tiles share their triangle records in L1); near the end of the queue single tiles keep the tail balanced.  The next
    // claim is issued while the run's last tile is processed, so its round trip is hidden.
    const int numWarps = int(gridDim.x) * 4;
    auto runLength = [&](int at) { return at + 8 * numWarps < totalTiles ? kTileRun : 1; };
    int run = runLength(0);
    int gw = 0;
    if (lane == 0) gw = atomicAdd(P.tileCounter, run);
    gw = __shfl_sync(0xffffffffu, gw, 0);
    int gwEnd = gw + run;
    for (;;) {
        if (gw >= totalTiles) return;
        const bool lastOfRun = gw + 1 >= gwEnd;
        const int nextRun = runLength(gw);
        int gwNext = gw + 1;
        if (lastOfRun && lane == 0) gwNext = atomicAdd(P.tileCounter, nextRun);
        const int vslot = gw / nTiles, tile = gw - vslot * nTiles;
        const int view = P.viewBase + vslot;
        const TriCover *cover = P.cover + size_t(vslot) * P.triCap;
        const TriShade *shade = P.shade + size_t(vslot) * P.triCap;
        int32_t *binCount = P.binCounts + size_t(vslot) * nTiles + tile;
        const uint16_t *bin = P.binList + (size_t(vslot) * nTiles + tile) * P.binCap;
        const int4 *wide = P.wideList + size_t(vslot) * kWideCap;
        const int nWide = min(__ldg(P.wideCounts + view), kWideCap);
        const int nBin = min(*binCount, P.binCap);
        const int total = nWide + nBin;

        const int ty = tile / tilesX;
        const int tx0 = (tile - ty * tilesX) * 32, ty0 = ty * 4;
        const int px = tx0 + (lane & 7) * 4, py = ty0 + (lane >> 3);
        const int sx32 = px * 256 + 128, sy32 = py * 256 + 128;
        unsigned long long best[4] = {0ull, 0ull, 0ull, 0ull};
        const long long tp0 = P.tileProf ? clock64() : 0;
        long long tpA = 0, tpB = 0;
        int nOv = 0, nSm = 0, nBg = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) frag[lane * 4 + k] = 0ull;
        __syncwarp();
        if (lane == 0 && nBin) *binCount = 0;  // leave the bin empty for the next geometry pass

        if (P.tileProf) tpA = clock64() + (total < 0 ? 1 : 0);  // (after the counts have arrived: total is consumed)
        for (int base = 0; base < total; base += 32) {
            const int j = base + lane;
            // which triangle is mine, and does its box touch this tile (bin entries always do; wide entries are checked)
            int tmine = -1;
            if (j < nWide) {
                const int4 w = __ldg(wide + j);
                const int x0 = w.x & 0xffff, x1 = int(unsigned(w.x) >> 16), y0 = w.y & 0xffff, y1 = int(unsigned(w.y) >> 16);
                if (x0 <= tx0 + 31 && x1 >= tx0 && y0 <= ty0 + 3 && y1 >= ty0) tmine = w.z;
            } else if (j < total) {
                tmine = int(bin[j - nWide]);
            }
            const bool ov = tmine >= 0;
            bool small = false;
            int bx0 = 0, bx1 = -1, by0 = 0, by1 = -1;
            EdgeEval e;
            e.small = 1;
            if (ov) {
                const int4 *cq = reinterpret_cast<const int4 *>(cover + tmine);
                const int4 q0 = __ldg(cq + 0), q1 = __ldg(cq + 1), q2 = __ldg(cq + 2), q3 = __ldg(cq + 3), q4 = __ldg(cq + 4);
                int4 *sq = reinterpret_cast<int4 *>(stage + lane);
                sq[0] = q0; sq[1] = q1; sq[2] = q2; sq[3] = q3; sq[4] = q4;
                e = unpackCover(q0, q1, q2, q3, q4);
                bx0 = max(int(unsigned(q4.z) & 0xffffu), tx0); bx1 = min(int(unsigned(q4.z) >> 16), tx0 + 31);
                by0 = max(int(unsigned(q4.w) & 0xffffu), ty0); by1 = min(int(unsigned(q4.w) >> 16), ty0 + 3);
                small = (bx1 - bx0 + 1) * (by1 - by0 + 1) <= kSmallArea;
            }
            if (P.tileProf) { nOv += __popc(__ballot_sync(0xffffffffu, ov)); nSm += __popc(__ballot_sync(0xffffffffu, small)); nBg += __popc(__ballot_sync(0xffffffffu, ov && !small)); }
            // ---- small triangles: one lane each
            if (small) {
                for (int y = by0; y <= by1; ++y) {
                    const int sy = y * 256 + 128, sx0 = bx0 * 256 + 128;
                    if (e.small) {
                        int F0 = int(e.C0) + e.A0 * sx0 + e.B0 * sy, F1 = int(e.C1) + e.A1 * sx0 + e.B1 * sy, F2 = int(e.C2) + e.A2 * sx0 + e.B2 * sy;
                        for (int x = bx0; x <= bx1; ++x) {
                            if ((F0 | F1 | F2) >= 0) {
                                const float l0 = float(F0 + e.u0) * e.invArea, l1 = float(F1 + e.u1) * e.invArea, l2 = float(F2 + e.u2) * e.invArea;
                                const float z = (l0 * e.z0 + l1 * e.z1) + l2 * e.z2;
                                if (z <= 1.0f) atomicMax(&frag[(y - ty0) * 32 + (x - tx0)], packFrag(z, e.key, tmine));
                            }
                            F0 += e.A0 * 256; F1 += e.A1 * 256; F2 += e.A2 * 256;
                        }
                    } else {
                        long long F0 = e.C0 + (long long)e.A0 * sx0 + (long long)e.B0 * sy, F1 = e.C1 + (long long)e.A1 * sx0 + (long long)e.B1 * sy,
                                  F2 = e.C2 + (long long)e.A2 * sx0 + (long long)e.B2 * sy;
                        for (int x = bx0; x <= bx1; ++x) {
                            if ((F0 | F1 | F2) >= 0) {
                                const float l0 = float(F0 + e.u0) * e.invArea, l1 = float(F1 + e.u1) * e.invArea, l2 = float(F2 + e.u2) * e.invArea;
                                const float z = (l0 * e.z0 + l1 * e.z1) + l2 * e.z2;
                                if (z <= 1.0f) atomicMax(&frag[(y - ty0) * 32 + (x - tx0)], packFrag(z, e.key, tmine));
                            }
                            F0 += (long long)e.A0 * 256; F1 += (long long)e.A1 * 256; F2 += (long long)e.A2 * 256;
                        }
                    }
                }
            }
            // ---- large triangles: whole warp, lane = 4 pixels; the record comes back from shared memory
            unsigned bits = __ballot_sync(0xffffffffu, ov && !small);
            __syncwarp();
            while (bits) {
                const int bsel = __ffs(bits) - 1;
                bits &= bits - 1;
                const int ti = __shfl_sync(0xffffffffu, tmine, bsel);
                const TriCover *sc = stage + bsel;
                const int x0 = int(sc->bx & 0xffffu), x1 = int(sc->bx >> 16), y0 = int(sc->by & 0xffffu), y1 = int(sc->by >> 16);
                if (px + 3 < x0 || px > x1 || py < y0 || py > y1) continue;
                const EdgeEval e2 = loadCoverShared(sc);
                if (e2.small) {
                    int F0 = int(e2.C0) + e2.A0 * sx32 + e2.B0 * sy32, F1 = int(e2.C1) + e2.A1 * sx32 + e2.B1 * sy32, F2 = int(e2.C2) + e2.A2 * sx32 + e2.B2 * sy32;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if ((F0 | F1 | F2) >= 0) {
                            const float l0 = float(F0 + e2.u0) * e2.invArea, l1 = float(F1 + e2.u1) * e2.invArea, l2 = float(F2 + e2.u2) * e2.invArea;
                            const float z = (l0 * e2.z0 + l1 * e2.z1) + l2 * e2.z2;
                            if (z <= 1.0f) { const unsigned long long f = packFrag(z, e2.key, ti); best[k] = f > best[k] ? f : best[k]; }
                        }
                        F0 += e2.A0 * 256; F1 += e2.A1 * 256; F2 += e2.A2 * 256;
                    }
                } else {
                    long long F0 = e2.C0 + (long long)e2.A0 * sx32 + (long long)e2.B0 * sy32, F1 = e2.C1 + (long long)e2.A1 * sx32 + (long long)e2.B1 * sy32,
                              F2 = e2.C2 + (long long)e2.A2 * sx32 + (long long)e2.B2 * sy32;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if ((F0 | F1 | F2) >= 0) {
                            const float l0 = float(F0 + e2.u0) * e2.invArea, l1 = float(F1 + e2.u1) * e2.invArea, l2 = float(F2 + e2.u2) * e2.invArea;
                            const float z = (l0 * e2.z0 + l1 * e2.z1) + l2 * e2.z2;
                            if (z <= 1.0f) { const unsigned long long f = packFrag(z, e2.key, ti); best[k] = f > best[k] ? f : best[k]; }
                        }
                        F0 += (long long)e2.A0 * 256; F1 += (long long)e2.A1 * 256; F2 += (long long)e2.A2 * 256;
                    }
                }
            }
            __syncwarp();  // the stage is rewritten by the next chunk
        }
        __syncwarp();
        if (P.tileProf) tpB = clock64();
        // ---- merge both paths, recompute the winner's barycentrics, shade, store
        uint4 out;
        float wv[4];
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned long long fs = frag[(lane >> 3) * 32 + (lane & 7) * 4 + k];
            const unsigned long long f = fs > best[k] ? fs : best[k];
            if (f == 0ull) { o[k] = 0xff000000u; wv[k] = 0.0f; continue; }
            const int ti = int(uint32_t(f) & 8191u);
            const ShadeRec rec = loadShade(shade + ti);  // both records are in flight together
            const EdgeEval e = loadCover(cover + ti);
            const int sx = sx32 + k * 256;
            float l0, l1, l2;
            if (e.small) {
                l0 = float(int(e.C0) + e.A0 * sx + e.B0 * sy32 + e.u0) * e.invArea;
                l1 = float(int(e.C1) + e.A1 * sx + e.B1 * sy32 + e.u1) * e.invArea;
                l2 = float(int(e.C2) + e.A2 * sx + e.B2 * sy32 + e.u2) * e.invArea;
            } else {
                l0 = float(e.C0 + (long long)e.A0 * sx + (long long)e.B0 * sy32 + e.u0) * e.invArea;
                l1 = float(e.C1 + (long long)e.A1 * sx + (long long)e.B1 * sy32 + e.u1) * e.invArea;
                l2 = float(e.C2 + (long long)e.A2 * sx + (long long)e.B2 * sy32 + e.u2) * e.invArea;
            }
            o[k] = shadePixel<FAST>(rec, l0, l1, l2, wv[k]);
        }
        out.x = o[0]; out.y = o[1]; out.z = o[2]; out.w = o[3];
        uint8_t *obsView = P.obs + size_t(view) * P.W * P.H * 4;
        *reinterpret_cast<uint4 *>(obsView + (size_t(py) * P.W + px) * 4) = out;
        if (P.depth) *reinterpret_cast<float4 *>(P.depth + size_t(view) * P.W * P.H + size_t(py) * P.W + px) = make_float4(wv[0], wv[1], wv[2], wv[3]);
        __syncwarp();
        if (P.sliceDone) {  // publish this tile's pixels, then count it
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(P.sliceDone + view / P.sliceViews, 1u);
        }
        if (P.tileProf && lane == 0) {
            uint32_t *tp = P.tileProf + (size_t(view) * nTiles + tile) * 4;
            tp[0] = uint32_t(clock64() - tp0); tp[1] = uint32_t(nOv) | (uint32_t(tpA - tp0) << 12); tp[2] = uint32_t(nSm) | (uint32_t(tpB - tpA) << 12); tp[3] = uint32_t(nBg);
        }
        if (lastOfRun) { gw = __shfl_sync(0xffffffffu, gwNext, 0); gwEnd = gw + nextRun; }
        else gw = gwNext;
    }
}

}  // namespace mvr
