// ORACLE (test infrastructure, NOT product code).  CPU restatement of the float math the
// reference's hot path goes through.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Every routine mirrors the *operation order* of the reference's math library so that float
// rounding matches:
//   Magnum  Math::RectangularMatrix::operator*   src/3rdparty/magnum/src/Magnum/Math/RectangularMatrix.h:753-764
//   Magnum  Math::Matrix::inverted (adjugate/det) src/3rdparty/magnum/src/Magnum/Math/Matrix.h:379-421,491-522
//   Magnum  Matrix4::rotation(angle, axis)        src/3rdparty/magnum/src/Magnum/Math/Matrix4.h:959-989
//   Magnum  Matrix4::rotationX                    src/3rdparty/magnum/src/Magnum/Math/Matrix4.h:991-999
//   Bullet  btQuaternion(axis,angle), btMatrix3x3::setRotation/getRotation, btQuaternion::getAxis/getAngle
//           -- Bullet 2.89 is NOT vendored in /root/reference; restated from the published upstream
//           algorithm (parity unpinned for these few routines; the stand-in Bullet the reference's env library is
//           compiled against for the pins forwards to them, see ref_shim/mini_bullet/mini_bullet.hpp and DESIGN.md section 6).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

// Transcendentals are evaluated in double and rounded once to float ("correctly rounded" up to ~1e-9 of inputs).
// The reference calls glibc's float sinf/cosf/acosf (Magnum: std::sin(Float); Bullet: btSin/btAcos -> sinf/acosf),
// which are NOT correctly rounded (measured here: 2-8 % of inputs differ by 1 ulp from the rounded double result) and
// differ between glibc versions, so they cannot be reproduced bit-for-bit on a GPU.  Oracle and device both use
// float(fn(double(x))): the reference's value is within 1 ulp of it.
inline float crsin(float x) { return float(std::sin(double(x))); }
inline float crcos(float x) { return float(std::cos(double(x))); }
inline float cracos(float x) { return float(std::acos(double(x))); }
inline float crtan(float x) { return float(std::tan(double(x))); }

struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, Vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 &operator+=(Vec3 &a, Vec3 b) { a = a + b; return a; }
inline Vec3 &operator-=(Vec3 &a, Vec3 b) { a = a - b; return a; }
inline Vec3 &operator*=(Vec3 &a, float s) { a = a * s; return a; }
// left-to-right sum, as both Magnum's Vector::dot and Bullet's btVector3::dot do
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length2(Vec3 a) { return dot(a, a); }
inline float length(Vec3 a) { return sqrtf(dot(a, a)); }
// btVector3::normalize(): *this /= length()  ==  *this * (1/length)   (btVector3 operator/=)
inline Vec3 btNormalized(Vec3 a) { float l = length(a); return a * (1.0f / l); }
// Magnum Vector::normalized(): *this * (1/sqrt(dot))  (Math/Vector.h: lengthInverted)
inline Vec3 mgNormalized(Vec3 a) { float l = 1.0f / sqrtf(dot(a, a)); return a * l; }

struct Mat3 {  // Bullet btMatrix3x3: three ROWS
    Vec3 r[3];
};
inline Mat3 mat3Identity() { Mat3 m; m.r[0] = {1, 0, 0}; m.r[1] = {0, 1, 0}; m.r[2] = {0, 0, 1}; return m; }
// btMatrix3x3 operator*: m[i].dot(column j of other) via tdotx/y/z
inline Mat3 mul(const Mat3 &a, const Mat3 &b) {
    Mat3 o;
    for (int i = 0; i < 3; ++i) {
        o.r[i].x = a.r[i].x * b.r[0].x + a.r[i].y * b.r[1].x + a.r[i].z * b.r[2].x;
        o.r[i].y = a.r[i].x * b.r[0].y + a.r[i].y * b.r[1].y + a.r[i].z * b.r[2].y;
        o.r[i].z = a.r[i].x * b.r[0].z + a.r[i].y * b.r[1].z + a.r[i].z * b.r[2].z;
    }
    return o;
}

struct Quat { float x = 0, y = 0, z = 0, w = 1; };
// btQuaternion(axis, angle)::setRotation
inline Quat quatAxisAngle(Vec3 axis, float angle) {
    float d = length(axis);
    float s = crsin(angle * 0.5f) / d;
    return {axis.x * s, axis.y * s, axis.z * s, crcos(angle * 0.5f)};
}
// btMatrix3x3::setRotation(q)
inline Mat3 mat3FromQuat(const Quat &q) {
    float d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    float s = 2.0f / d;
    float xs = q.x * s, ys = q.y * s, zs = q.z * s;
    float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    Mat3 m;
    m.r[0] = {1.0f - (yy + zz), xy - wz, xz + wy};
    m.r[1] = {xy + wz, 1.0f - (xx + zz), yz - wx};
    m.r[2] = {xz - wy, yz + wx, 1.0f - (xx + yy)};
    return m;
}
// btMatrix3x3::getRotation(q) (scalar path)
inline Quat quatFromMat3(const Mat3 &m) {
    float trace = m.r[0].x + m.r[1].y + m.r[2].z;
    float t[4];
    if (trace > 0.0f) {
        float s = sqrtf(trace + 1.0f);
        t[3] = s * 0.5f;
        s = 0.5f / s;
        t[0] = (m.r[2].y - m.r[1].z) * s;
        t[1] = (m.r[0].z - m.r[2].x) * s;
        t[2] = (m.r[1].x - m.r[0].y) * s;
    } else {
        int i = m.r[0].x < m.r[1].y ? (m.r[1].y < m.r[2].z ? 2 : 1) : (m.r[0].x < m.r[2].z ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        float s = sqrtf(m.r[i][i] - m.r[j][j] - m.r[k][k] + 1.0f);
        t[i] = s * 0.5f;
        s = 0.5f / s;
        t[3] = (m.r[k][j] - m.r[j][k]) * s;
        t[j] = (m.r[j][i] + m.r[i][j]) * s;
        t[k] = (m.r[k][i] + m.r[i][k]) * s;
    }
    return {t[0], t[1], t[2], t[3]};
}
// btQuaternion::getAngle: 2*acos(w);  getAxis: s2 = 1-w*w; if (s2 < 10*eps) (1,0,0) else xyz/sqrt(s2)
// (btAcos clamps its argument to [-1, 1], btScalar.h: a w that rounds above 1 gives angle 0, not NaN)
inline float quatAngle(const Quat &q) { float w = q.w; if (w < -1.0f) w = -1.0f; if (w > 1.0f) w = 1.0f; return 2.0f * cracos(w); }
inline Vec3 quatAxis(const Quat &q) {
    float s2 = 1.0f - q.w * q.w;
    if (s2 < 10.0f * 1.1920929e-07f) return {1, 0, 0};
    float s = 1.0f / sqrtf(s2);
    return {q.x * s, q.y * s, q.z * s};
}

struct Mat4 {  // Magnum Matrix4: c[col][row], column-major
    float c[4][4];
};
inline Mat4 mat4Identity() {
    Mat4 m;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m.c[i][j] = i == j ? 1.0f : 0.0f;
    return m;
}
// RectangularMatrix::operator*: out zero-init, += this[pos][row]*other[col][pos], pos ascending
inline Mat4 mul(const Mat4 &a, const Mat4 &b) {
    Mat4 o;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            float acc = 0.0f;
            for (int pos = 0; pos < 4; ++pos) acc += a.c[pos][row] * b.c[col][pos];
            o.c[col][row] = acc;
        }
    return o;
}
inline Mat4 mat4Translation(Vec3 t) { Mat4 m = mat4Identity(); m.c[3][0] = t.x; m.c[3][1] = t.y; m.c[3][2] = t.z; return m; }
inline Mat4 mat4Scaling(Vec3 s) { Mat4 m = mat4Identity(); m.c[0][0] = s.x; m.c[1][1] = s.y; m.c[2][2] = s.z; return m; }
inline Mat4 mat4RotationX(float a) {
    float s = crsin(a), c = crcos(a);
    Mat4 m = mat4Identity();
    m.c[1][1] = c; m.c[1][2] = s; m.c[2][1] = -s; m.c[2][2] = c;
    return m;
}
inline Mat4 mat4RotationY(float a) {
    float s = crsin(a), c = crcos(a);
    Mat4 m = mat4Identity();
    m.c[0][0] = c; m.c[0][2] = -s; m.c[2][0] = s; m.c[2][2] = c;
    return m;
}
inline Mat4 mat4Rotation(float angle, Vec3 ax) {
    float sine = crsin(angle), cosine = crcos(angle), omc = 1.0f - cosine;
    float xx = ax.x * ax.x, xy = ax.x * ax.y, xz = ax.x * ax.z, yy = ax.y * ax.y, yz = ax.y * ax.z, zz = ax.z * ax.z;
    Mat4 m = mat4Identity();
    m.c[0][0] = cosine + xx * omc; m.c[0][1] = xy * omc + ax.z * sine; m.c[0][2] = xz * omc - ax.y * sine;
    m.c[1][0] = xy * omc - ax.z * sine; m.c[1][1] = cosine + yy * omc; m.c[1][2] = yz * omc + ax.x * sine;
    m.c[2][0] = xz * omc + ax.y * sine; m.c[2][1] = yz * omc - ax.x * sine; m.c[2][2] = cosine + zz * omc;
    return m;
}
inline Vec3 translationOf(const Mat4 &m) { return {m.c[3][0], m.c[3][1], m.c[3][2]}; }
// Matrix4::scaling(): lengths of the three basis columns
inline Vec3 scalingOf(const Mat4 &m) {
    return {sqrtf(m.c[0][0] * m.c[0][0] + m.c[0][1] * m.c[0][1] + m.c[0][2] * m.c[0][2]),
            sqrtf(m.c[1][0] * m.c[1][0] + m.c[1][1] * m.c[1][1] + m.c[1][2] * m.c[1][2]),
            sqrtf(m.c[2][0] * m.c[2][0] + m.c[2][1] * m.c[2][1] + m.c[2][2] * m.c[2][2])};
}
// Matrix4::transformPoint: (M * vec4(p,1)).xyz, same accumulation order as operator*
inline Vec3 transformPoint(const Mat4 &m, Vec3 p) {
    Vec3 o;
    for (int row = 0; row < 3; ++row) {
        float acc = 0.0f;
        acc += m.c[0][row] * p.x; acc += m.c[1][row] * p.y; acc += m.c[2][row] * p.z; acc += m.c[3][row] * 1.0f;
        o[row] = acc;
    }
    return o;
}
// MatrixDeterminant<3>(Matrix<4>, skipCol, skipRow)
inline float det3skip(const Mat4 &m, int skipCol, int skipRow) {
    auto C = [&](int i) { return i + (i >= skipCol); };
    auto R = [&](int i) { return i + (i >= skipRow); };
    return m.c[C(0)][R(0)] * ((m.c[C(1)][R(1)] * m.c[C(2)][R(2)]) - (m.c[C(2)][R(1)] * m.c[C(1)][R(2)])) -
           m.c[C(0)][R(1)] * (m.c[C(1)][R(0)] * m.c[C(2)][R(2)] - m.c[C(2)][R(0)] * m.c[C(1)][R(2)]) +
           m.c[C(0)][R(2)] * (m.c[C(1)][R(0)] * m.c[C(2)][R(1)] - m.c[C(2)][R(0)] * m.c[C(1)][R(1)]);
}
inline float cofactor4(const Mat4 &m, int col, int row) { return (((row + col) & 1) ? -1 : 1) * det3skip(m, col, row); }
inline float det4(const Mat4 &m) {
    float out = 0.0f;
    for (int col = 0; col < 4; ++col) out += m.c[col][0] * cofactor4(m, col, 0);
    return out;
}
inline Mat4 inverted(const Mat4 &m) {
    Mat4 o;
    float d = det4(m);
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) o.c[col][row] = cofactor4(m, row, col) / d;
    return o;
}
// 3x3 inverse-transpose for normals: GLSL transpose(inverse(mat3(mv))) -- the GLSL compiler's
// expansion is not observable; define as cofactor matrix / determinant (inverse-transpose == comatrix/det).
inline void normalMatrix(const Mat4 &mv, float n[3][3]) {  // n[col][row]
    const float a00 = mv.c[0][0], a01 = mv.c[0][1], a02 = mv.c[0][2];
    const float a10 = mv.c[1][0], a11 = mv.c[1][1], a12 = mv.c[1][2];
    const float a20 = mv.c[2][0], a21 = mv.c[2][1], a22 = mv.c[2][2];
    const float c00 = a11 * a22 - a21 * a12, c01 = a20 * a12 - a10 * a22, c02 = a10 * a21 - a20 * a11;
    const float c10 = a21 * a02 - a01 * a22, c11 = a00 * a22 - a20 * a02, c12 = a20 * a01 - a00 * a21;
    const float c20 = a01 * a12 - a11 * a02, c21 = a10 * a02 - a00 * a12, c22 = a00 * a11 - a10 * a01;
    const float det = a00 * c00 + a01 * c01 + a02 * c02;
    const float id = 1.0f / det;
    n[0][0] = c00 * id; n[0][1] = c01 * id; n[0][2] = c02 * id;
    n[1][0] = c10 * id; n[1][1] = c11 * id; n[1][2] = c12 * id;
    n[2][0] = c20 * id; n[2][1] = c21 * id; n[2][2] = c22 * id;
}

}  // namespace orc
