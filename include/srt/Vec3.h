// srt/Vec3.h -- 3-component float vector with the reference's public API
// (include/Vec3.h:138-201, include/Vec3.inl).  Semantics that matter for bit-parity are kept:
//  * float and double scalar overloads: the double ones multiply/divide in double and narrow
//    when the result is stored (Vec3.inl:221-233);
//  * dotProduct sums left to right (Vec3.inl:313-317); normalise divides each component by the
//    length (three divides, not a multiply by the reciprocal; Vec3.inl:469-476);
//  * operator[] throws std::out_of_range outside [0, 2] (Vec3.inl:333-363).
// Build host code that uses it for parity-sensitive work with -ffp-contract=off.
#pragma once
#include <cmath>
#include <ostream>
#include <stdexcept>

namespace srt {

class Vec3 {
public:
    Vec3(float x = 0.0f, float y = 0.0f, float z = 0.0f) : c_{ x, y, z } {}

    Vec3 operator-() const { return { -c_[0], -c_[1], -c_[2] }; }
    Vec3 operator+(const Vec3& o) const { return { c_[0] + o.c_[0], c_[1] + o.c_[1], c_[2] + o.c_[2] }; }
    Vec3 operator-(const Vec3& o) const { return { c_[0] - o.c_[0], c_[1] - o.c_[1], c_[2] - o.c_[2] }; }
    Vec3& operator+=(const Vec3& o) { for (int i = 0; i < 3; ++i) c_[i] += o.c_[i]; return *this; }
    Vec3& operator-=(const Vec3& o) { for (int i = 0; i < 3; ++i) c_[i] -= o.c_[i]; return *this; }

    Vec3 operator*(float t) const { return { c_[0] * t, c_[1] * t, c_[2] * t }; }
    Vec3 operator/(float t) const { return { c_[0] / t, c_[1] / t, c_[2] / t }; }
    // double overloads: arithmetic in double, narrowed by the constructor
    Vec3 operator*(double t) const { return { float(c_[0] * t), float(c_[1] * t), float(c_[2] * t) }; }
    Vec3 operator/(double t) const { return { float(c_[0] / t), float(c_[1] / t), float(c_[2] / t) }; }
    Vec3& operator*=(float t) { for (float& v : c_) v *= t; return *this; }
    Vec3& operator/=(float t) { for (float& v : c_) v /= t; return *this; }
    Vec3& operator*=(double t) { for (float& v : c_) v = float(v * t); return *this; }
    Vec3& operator/=(double t) { for (float& v : c_) v = float(v / t); return *this; }

    Vec3 operator*(const Vec3& o) const { return { c_[0] * o.c_[0], c_[1] * o.c_[1], c_[2] * o.c_[2] }; }
    Vec3& operator*=(const Vec3& o) { for (int i = 0; i < 3; ++i) c_[i] *= o.c_[i]; return *this; }

    float dotProduct(const Vec3& o) const { return c_[0] * o.c_[0] + c_[1] * o.c_[1] + c_[2] * o.c_[2]; }
    Vec3 crossProduct(const Vec3& o) const
    {
        return { c_[1] * o.c_[2] - c_[2] * o.c_[1], c_[2] * o.c_[0] - c_[0] * o.c_[2], c_[0] * o.c_[1] - c_[1] * o.c_[0] };
    }

    float& operator[](unsigned int i) { check(i); return c_[i]; }
    const float& operator[](unsigned int i) const { check(i); return c_[i]; }

    float getX() const { return c_[0]; }
    float getY() const { return c_[1]; }
    float getZ() const { return c_[2]; }
    float getR() const { return c_[0]; }
    float getG() const { return c_[1]; }
    float getB() const { return c_[2]; }
    void setX(float v) { c_[0] = v; }
    void setY(float v) { c_[1] = v; }
    void setZ(float v) { c_[2] = v; }
    void setR(float v) { c_[0] = v; }
    void setG(float v) { c_[1] = v; }
    void setB(float v) { c_[2] = v; }

    float getLength() const { return std::sqrt(c_[0] * c_[0] + c_[1] * c_[1] + c_[2] * c_[2]); }
    void normalise() { const float len = getLength(); c_[0] /= len; c_[1] /= len; c_[2] /= len; }
    void normalize() { normalise(); }

    const float* data() const { return c_; }          // extension: contiguous access for the C ABI

private:
    static void check(unsigned int i) { if (i > 2) throw std::out_of_range("Valid range is [0, 2]"); }
    float c_[3];
};

inline Vec3 normalise(const Vec3& v) { Vec3 t(v); t.normalise(); return t; }
inline Vec3 normalize(const Vec3& v) { return normalise(v); }
inline Vec3 operator*(float t, const Vec3& v) { return v * t; }
inline Vec3 operator*(double t, const Vec3& v) { return v * t; }
inline std::ostream& operator<<(std::ostream& os, const Vec3& v) { return os << v.getX() << " " << v.getY() << " " << v.getZ(); }
// reflect(I, N) = I - 2.0 * N.I * N : the 2.0 makes the scalar a double (Vec3.inl:120-124)
inline Vec3 reflect(const Vec3& I, const Vec3& N) { return I - 2.0 * N.dotProduct(I) * N; }
inline float dot(const Vec3& a, const Vec3& b) { return a.dotProduct(b); }

} // namespace srt
