// Host-side float math for level generation (product code).  Column-major 4x4 products keep the accumulation order of
// Magnum's RectangularMatrix::operator* (src/3rdparty/magnum/src/Magnum/Math/RectangularMatrix.h:753-764) so model
// matrices built here carry the same bits the reference's scene graph would produce.  Transcendentals are evaluated in
// double and rounded once (see DESIGN.md "numerics").
#pragma once
#include <cmath>
#include <cstring>

namespace mvh {

inline float crsin(float x) { return float(std::sin(double(x))); }
inline float crcos(float x) { return float(std::cos(double(x))); }
inline float crtan(float x) { return float(std::tan(double(x))); }

struct M4 { float c[4][4]; };
inline M4 identity() { M4 m; std::memset(&m, 0, sizeof m); m.c[0][0] = m.c[1][1] = m.c[2][2] = m.c[3][3] = 1.0f; return m; }
inline M4 mul(const M4 &a, const M4 &b) {
    M4 o;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            float acc = 0.0f;
            for (int pos = 0; pos < 4; ++pos) acc += a.c[pos][row] * b.c[col][pos];
            o.c[col][row] = acc;
        }
    return o;
}
inline M4 translation(float x, float y, float z) { M4 m = identity(); m.c[3][0] = x; m.c[3][1] = y; m.c[3][2] = z; return m; }
inline M4 scaling(float x, float y, float z) { M4 m = identity(); m.c[0][0] = x; m.c[1][1] = y; m.c[2][2] = z; return m; }
// Magnum Matrix::inverted(): adjugate / determinant, same operation order as dev_math.cuh's inverted4 and the oracle's
inline float det3skip(const M4 &m, int skipCol, int skipRow) {
#define MVH_E(ci, ri) m.c[(ci) + ((ci) >= skipCol)][(ri) + ((ri) >= skipRow)]
    return MVH_E(0, 0) * ((MVH_E(1, 1) * MVH_E(2, 2)) - (MVH_E(2, 1) * MVH_E(1, 2))) - MVH_E(0, 1) * (MVH_E(1, 0) * MVH_E(2, 2) - MVH_E(2, 0) * MVH_E(1, 2)) +
           MVH_E(0, 2) * (MVH_E(1, 0) * MVH_E(2, 1) - MVH_E(2, 0) * MVH_E(1, 1));
#undef MVH_E
}
inline float cofactor4(const M4 &m, int col, int row) { return (((row + col) & 1) ? -1 : 1) * det3skip(m, col, row); }
inline M4 inverted(const M4 &m) {
    float d = 0.0f;
    for (int col = 0; col < 4; ++col) d += m.c[col][0] * cofactor4(m, col, 0);
    M4 o;
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) o.c[col][row] = cofactor4(m, row, col) / d;
    return o;
}
inline M4 rotationY(float a) {
    const float s = crsin(a), c = crcos(a);
    M4 m = identity();
    m.c[0][0] = c; m.c[0][2] = -s; m.c[2][0] = s; m.c[2][2] = c;
    return m;
}

// rows of btMatrix3x3(btQuaternion(Y axis, angle))  -- agent.cpp:44,128-133
inline void yawBasis(float angle, float rows[9]) {
    const float d = sqrtf(0.0f * 0.0f + 1.0f * 1.0f + 0.0f * 0.0f);
    const float s = crsin(angle * 0.5f) / d;
    const float qx = 0.0f * s, qy = 1.0f * s, qz = 0.0f * s, qw = crcos(angle * 0.5f);
    const float dd = qx * qx + qy * qy + qz * qz + qw * qw;
    const float k = 2.0f / dd;
    const float xs = qx * k, ys = qy * k, zs = qz * k;
    const float wx = qw * xs, wy = qw * ys, wz = qw * zs;
    const float xx = qx * xs, xy = qx * ys, xz = qx * zs;
    const float yy = qy * ys, yz = qy * zs, zz = qz * zs;
    rows[0] = 1.0f - (yy + zz); rows[1] = xy - wz; rows[2] = xz + wy;
    rows[3] = xy + wz; rows[4] = 1.0f - (xx + zz); rows[5] = yz - wx;
    rows[6] = xz - wy; rows[7] = yz + wx; rows[8] = 1.0f - (xx + yy);
}

// btMatrix3x3::setRotation(q) / getRotation(q) (scalar paths) on row-major 3x3s
inline void rowsFromQuat(const float q[4], float rows[9]) {
    const float dd = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const float k = 2.0f / dd;
    const float xs = q[0] * k, ys = q[1] * k, zs = q[2] * k;
    const float wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    const float xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
    const float yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    rows[0] = 1.0f - (yy + zz); rows[1] = xy - wz; rows[2] = xz + wy;
    rows[3] = xy + wz; rows[4] = 1.0f - (xx + zz); rows[5] = yz - wx;
    rows[6] = xz - wy; rows[7] = yz + wx; rows[8] = 1.0f - (xx + yy);
}
inline void quatFromRows(const float m[9], float q[4]) {
    const float trace = m[0] + m[4] + m[8];
    if (trace > 0.0f) {
        float s = sqrtf(trace + 1.0f);
        q[3] = s * 0.5f;
        s = 0.5f / s;
        q[0] = (m[7] - m[5]) * s;
        q[1] = (m[2] - m[6]) * s;
        q[2] = (m[3] - m[1]) * s;
    } else {
        const int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        float s = sqrtf(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0f);
        q[i] = s * 0.5f;
        s = 0.5f / s;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * s;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * s;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * s;
    }
}
// the ghost basis a freshly spawned agent ends up with: btQuaternion(Y, angle) -> matrix (agent.cpp:42-44), then the character
// controller's constructor (setUp -> setGravity -> setUpVector, kinematic_character_controller.cpp:148,646-651,727-749) reads the
// rotation back as a quaternion, multiplies by the identity correction and stores the matrix again -- not always a no-op in float
inline void spawnBasis(float angle, float rows[9]) {
    float first[9], q[4];
    yawBasis(angle, first);
    quatFromRows(first, q);
    rowsFromQuat(q, rows);
}

}  // namespace mvh
