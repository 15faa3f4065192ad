// srt/Triangle.h -- triangle with cached unit normal and texture coordinates
// (reference API: include/Triangle.h:92-134, include/Triangle.inl).
#pragma once
#include <cmath>
#include <ostream>

#include "Vec3.h"

namespace srt {

class Triangle {
public:
    Triangle(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& ta = Vec3(), const Vec3& tb = Vec3(),
             const Vec3& tc = Vec3())
        : p_{ a, b, c }, tex_{ ta, tb, tc } { computeNormal(); }

    const Vec3& getP1() const { return p_[0]; }
    const Vec3& getP2() const { return p_[1]; }
    const Vec3& getP3() const { return p_[2]; }
    const Vec3& getNormal() const { return normal_; }
    const Vec3& getTextCoord1() const { return tex_[0]; }
    const Vec3& getTextCoord2() const { return tex_[1]; }
    const Vec3& getTextCoord3() const { return tex_[2]; }

    void setVertices(const Vec3& a, const Vec3& b, const Vec3& c) { p_[0] = a; p_[1] = b; p_[2] = c; computeNormal(); }
    void setTextCoords(const Vec3& a, const Vec3& b, const Vec3& c) { tex_[0] = a; tex_[1] = b; tex_[2] = c; }

    // 0.5 * sqrt(|AB x AC|^2) with the cross components in float and pow/sqrt in double
    // (Triangle.inl:206-216)
    float getArea() const
    {
        const Vec3 ab = p_[1] - p_[0], ac = p_[2] - p_[0];
        const double x = ab[1] * ac[2] - ab[2] * ac[1], y = ab[2] * ac[0] - ab[0] * ac[2], z = ab[0] * ac[1] - ab[1] * ac[0];
        return float(0.5 * std::sqrt(std::pow(x, 2) + std::pow(y, 2) + std::pow(z, 2)));
    }

private:
    // unit normal of (P2 - P1) x (P3 - P1) (Triangle.inl:170-178)
    void computeNormal() { normal_ = (p_[1] - p_[0]).crossProduct(p_[2] - p_[0]); normal_.normalise(); }

    Vec3 p_[3];
    Vec3 normal_;
    Vec3 tex_[3];
};

// STL-like text dump, same layout as the reference's operator<< (src/Triangle.cxx:77-88)
inline std::ostream& operator<<(std::ostream& os, const Triangle& t)
{
    os << "facet normal " << t.getNormal() << "\nouter loop\n";
    for (const Vec3* p : { &t.getP1(), &t.getP2(), &t.getP3() }) os << "    vertex " << *p << "\n";
    return os << "endloop\nendfacet";
}

} // namespace srt
