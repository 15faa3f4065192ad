// srt/Ray.h -- ray with the reference's API (include/Ray.h:84-107, include/Ray.inl, src/Ray.cxx).
//
// The constructor re-normalises the direction unless its length classifies as zero (Ray.inl:74-85).
// Ray::intersect is the reference's single-ray Moller-Trumbore (src/Ray.cxx:72-125) kept as a HOST
// UTILITY for callers that probe individual rays; the render entry points (srt/RenderLoop.h) never
// call it -- frames are rendered by the CUDA kernels behind include/srt.h, and fail if no GPU is
// present.
#pragma once
#include <cmath>

#include "Triangle.h"
#include "Vec3.h"

namespace srt {

class Ray {
public:
    Ray(const Vec3& origin, const Vec3& direction) : origin_(origin)
    {
        const float len = direction.getLength();
        if (std::fpclassify(len) != FP_ZERO) direction_ = direction / len;
    }
    Ray(const Ray&) = default;               // declared but never defined in the reference (Ray.h:88)
    Ray& operator=(const Ray&) = default;

    const Vec3& getOrigin() const { return origin_; }
    const Vec3& getDirection() const { return direction_; }
    void setOrigin(const Vec3& p) { origin_ = p; }
    void setDirection(const Vec3& d) { direction_ = d / d.getLength(); }
    Vec3 getPointAt(float t) const { return origin_ + direction_ * t; }

    bool intersect(const Triangle& tri, float& t) const
    {
        const Vec3 e1 = tri.getP2() - tri.getP1(), e2 = tri.getP3() - tri.getP1();
        const Vec3 pvec = direction_.crossProduct(e2);
        const float det = e1.dotProduct(pvec);
        if (std::fpclassify(det) == FP_ZERO) return false;
        const float inv = float(1.0 / det);
        const Vec3 tvec = origin_ - tri.getP1();
        const float u = tvec.dotProduct(pvec) * inv;
        if (u < 0.0 || u > 1.0) return false;
        const Vec3 qvec = tvec.crossProduct(e1);
        const float v = direction_.dotProduct(qvec) * inv;
        if (v < 0.0 || u + v > 1.0) return false;
        t = e2.dotProduct(qvec) * inv;
        return true;
    }

private:
    Vec3 origin_;
    Vec3 direction_;
};

} // namespace srt
