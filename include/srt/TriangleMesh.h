// srt/TriangleMesh.h -- triangle soup + material + bounding box + texture, the reference's API
// (include/TriangleMesh.h:106-169, src/TriangleMesh.cxx).  Five setGeometry overloads; sizes that are
// not multiples of 3 / 9 (or mismatching texture-coordinate sets) throw std::length_error like the
// reference (src/TriangleMesh.cxx:98,129,157,186).  intersectBBox is the reference's stub: always true
// (include/TriangleMesh.inl:232-236).
#pragma once
#include <algorithm>
#include <limits>
#include <ostream>
#include <stdexcept>
#include <vector>

#include "Image.h"
#include "ImageRGB.h"
#include "Material.h"
#include "Ray.h"
#include "Triangle.h"

namespace srt {

class TriangleMesh {
public:
    TriangleMesh() = default;
    explicit TriangleMesh(const std::vector<float>& vertices) { setGeometry(vertices); }
    TriangleMesh(const std::vector<float>& vertices, const std::vector<unsigned int>& indices) { setGeometry(vertices, indices); }
    TriangleMesh(const std::vector<float>& vertices, const std::vector<float>& texcoords) { setGeometry(vertices, texcoords); }
    TriangleMesh(const std::vector<float>& vertices, const std::vector<unsigned int>& indices,
                 const std::vector<float>& texcoords) { setGeometry(vertices, indices, texcoords); }
    explicit TriangleMesh(const std::vector<Triangle>& triangles) { setGeometry(triangles); }

    // 9 floats per triangle
    void setGeometry(const std::vector<float>& v)
    {
        if (v.size() % 9 != 0) throw std::length_error("buffer size error");
        triangles_.clear();
        triangles_.reserve(v.size() / 9);
        for (size_t k = 0; k < v.size(); k += 9) triangles_.emplace_back(vec(v, k / 3), vec(v, k / 3 + 1), vec(v, k / 3 + 2));
        computeBoundingBox();
    }
    // indexed: 3 floats per vertex, 3 indices per face
    void setGeometry(const std::vector<float>& v, const std::vector<unsigned int>& idx)
    {
        if (v.size() % 3 != 0 || idx.size() % 3 != 0) throw std::length_error("buffer size error");
        triangles_.clear();
        triangles_.reserve(idx.size() / 3);
        for (size_t f = 0; f < idx.size(); f += 3) triangles_.emplace_back(vec(v, idx[f]), vec(v, idx[f + 1]), vec(v, idx[f + 2]));
        computeBoundingBox();
    }
    void setGeometry(const std::vector<float>& v, const std::vector<float>& tc)
    {
        setGeometry(v);
        if (v.size() != tc.size()) throw std::length_error("buffer size error");
        for (size_t t = 0; t < triangles_.size(); ++t) triangles_[t].setTextCoords(vec(tc, 3 * t), vec(tc, 3 * t + 1), vec(tc, 3 * t + 2));
    }
    void setGeometry(const std::vector<float>& v, const std::vector<unsigned int>& idx, const std::vector<float>& tc)
    {
        setGeometry(v, idx);
        if (v.size() != tc.size()) throw std::length_error("buffer size error");
        for (size_t f = 0; f < idx.size(); f += 3) triangles_[f / 3].setTextCoords(vec(tc, idx[f]), vec(tc, idx[f + 1]), vec(tc, idx[f + 2]));
    }
    void setGeometry(const std::vector<Triangle>& triangles) { triangles_ = triangles; computeBoundingBox(); }

    void setMaterial(const Material& m) { material_ = m; }
    Material& getMaterial() { return material_; }
    const Material& getMaterial() const { return material_; }
    // The shaded drivers texture a mesh with the (older) RGB image class; the float Image overload of the current
    // reference header (TriangleMesh.h:150-152) is kept for source compatibility.
    void setTexture(const ImageRGB& img) { texture_rgb_ = img; }
    void setTexture(const Image& img) { texture_ = img; }
    Image& getTexture() { return texture_; }
    const Image& getTexture() const { return texture_; }
    const ImageRGB& getTextureRGB() const { return texture_rgb_; }
    // per-triangle texture coordinates, 9 floats each (extension for the C-ABI shim)
    std::vector<float> flatTexCoords() const
    {
        std::vector<float> out; out.reserve(9 * triangles_.size());
        for (const Triangle& t : triangles_)
            for (const Vec3* p : { &t.getTextCoord1(), &t.getTextCoord2(), &t.getTextCoord3() }) out.insert(out.end(), p->data(), p->data() + 3);
        return out;
    }

    size_t getNumberOfTriangles() const { return triangles_.size(); }
    const Triangle& getTriangle(unsigned int i) const { return triangles_[i]; }
    const Vec3& getLowerBBoxCorner() const { return lower_; }
    const Vec3& getUpperBBoxCorner() const { return upper_; }
    bool intersectBBox(const Ray&) const { return true; }

    // extension used by the C-ABI shim: flat 9-floats-per-triangle / 3-floats-per-normal copies
    std::vector<float> flatVertices() const
    {
        std::vector<float> out; out.reserve(9 * triangles_.size());
        for (const Triangle& t : triangles_)
            for (const Vec3* p : { &t.getP1(), &t.getP2(), &t.getP3() }) out.insert(out.end(), p->data(), p->data() + 3);
        return out;
    }
    std::vector<float> flatNormals() const
    {
        std::vector<float> out; out.reserve(3 * triangles_.size());
        for (const Triangle& t : triangles_) out.insert(out.end(), t.getNormal().data(), t.getNormal().data() + 3);
        return out;
    }

protected:
    // min / max over P1, P2, P3 of every triangle (src/TriangleMesh.cxx:192-228)
    void computeBoundingBox()
    {
        const float inf = std::numeric_limits<float>::infinity();
        lower_ = Vec3(inf, inf, inf); upper_ = Vec3(-inf, -inf, -inf);
        for (const Triangle& t : triangles_)
            for (const Vec3* p : { &t.getP1(), &t.getP2(), &t.getP3() })
                for (unsigned int k = 0; k < 3; ++k) {
                    lower_[k] = std::min(lower_[k], (*p)[k]);
                    upper_[k] = std::max(upper_[k], (*p)[k]);
                }
    }

private:
    static Vec3 vec(const std::vector<float>& v, size_t i) { return Vec3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); }

    std::vector<Triangle> triangles_;
    Material material_;
    Vec3 lower_, upper_;
    Image texture_;
    ImageRGB texture_rgb_;
};

inline std::ostream& operator<<(std::ostream& os, const TriangleMesh& m)
{
    os << "solid name\n";
    for (size_t i = 0; i < m.getNumberOfTriangles(); ++i) os << m.getTriangle((unsigned int)i) << "\n";
    return os << "endsolid name";
}

} // namespace srt
