// ORACLE tooling -- NOT product code.  Evaluates the reference's lit fragment stage FROM ITS OWN TEXT: the Makefile cuts
// `vec4 compute_color()` (the LIT_PIPELINE definition) out of src/3rdparty/v4r/src/pipelines/shaders/uber.frag at build time into
// oracle/_ref/include/v4r_compute_color.inc (a derived file in the git-ignored _ref directory; unsuffixed GLSL float literals get an
// `f`, which is what they mean in GLSL) and this file compiles it as C++ against the glm vendored with V4R (glm implements the GLSL
// built-ins normalize / dot / reflect / pow / clamp / max).  The environment the shader sees is the one the reference sets up:
// DIFFUSE / SPECULAR / SHININESS uniforms from the material (v4r_env_renderer.cpp:204-213), one light (:220), no V4R_BLINN_PHONG.
// The vertex stage's normal matrix (uber.vert:86-87: transpose(inverse(mat3(mv)))) is evaluated with the same glm calls.
#include <cstdint>
#include <cstring>

#define GLM_FORCE_SWIZZLE  // the shader text uses .rgb / .xyz
#include <glm/glm.hpp>

using namespace glm;

namespace {
struct Light { vec4 position, color; };
struct LightingInfo { Light lights[1]; int numLights; };
struct MaterialParams { vec3 diffuse, specular; float shininess; };

LightingInfo lighting_info{{{vec4(0.f, 4.f, 2.f, 1.f), vec4(0.66f, 0.66f, 0.66f, 1.f)}}, 1};  // v4r_env_renderer.cpp:220
MaterialParams material_params[1];
const unsigned material_idx = 0;
vec3 in_camera_pos, in_normal;

#define MATERIAL_PARAMS
#define DIFFUSE_COLOR_UNIFORM
#define DIFFUSE_COLOR_ACCESS params.diffuse
#define SPECULAR_COLOR_UNIFORM
#define SPECULAR_COLOR_ACCESS params.specular
#define SHININESS_UNIFORM
#define SHININESS_ACCESS params.shininess
#include <v4r_compute_color.inc>
}  // namespace

extern "C" {

void ref_shade(int n, const float *P3, const float *N3, const float *diffuse3, float *out3) {
    material_params[0].specular = vec3(1.f, 1.f, 1.f);  // v4r_env_renderer.cpp:204-213
    material_params[0].shininess = 300.f;
    for (int i = 0; i < n; ++i) {
        in_camera_pos = vec3(P3[3 * i], P3[3 * i + 1], P3[3 * i + 2]);
        in_normal = vec3(N3[3 * i], N3[3 * i + 1], N3[3 * i + 2]);
        material_params[0].diffuse = vec3(diffuse3[3 * i], diffuse3[3 * i + 1], diffuse3[3 * i + 2]);
        const vec4 c = compute_color();
        out3[3 * i] = c.x, out3[3 * i + 1] = c.y, out3[3 * i + 2] = c.z;
    }
}

void ref_normal_matrix(const float *mv16, float *out9) {  // uber.vert:86: mat3 normal_mat = transpose(inverse(mat3(mv)));
    mat4 mv;
    std::memcpy(&mv[0][0], mv16, 64);
    const mat3 normal_mat = transpose(inverse(mat3(mv)));
    std::memcpy(out9, &normal_mat[0][0], 36);
}

}  // extern "C"
