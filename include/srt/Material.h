// srt/Material.h, Light -- plain data used by the shaded variant
// (reference API: include/Material.h:86-124, include/Light.h:80-111).
#pragma once
#include "Vec3.h"

namespace srt {

class Material {
public:
    Material() : ambient_(1, 1, 1), diffuse_(1, 1, 1), specular_(1, 1, 1), shininess_(1) {}
    Material(const Vec3& ambient, const Vec3& diffuse, const Vec3& specular, float shininess)
        : ambient_(ambient), diffuse_(diffuse), specular_(specular), shininess_(shininess) {}
    void setAmbient(const Vec3& v) { ambient_ = v; }
    void setDiffuse(const Vec3& v) { diffuse_ = v; }
    void setSpecular(const Vec3& v) { specular_ = v; }
    void setShininess(float s) { shininess_ = s; }
    Vec3& getAmbient() { return ambient_; }
    const Vec3& getAmbient() const { return ambient_; }
    Vec3& getDiffuse() { return diffuse_; }
    const Vec3& getDiffuse() const { return diffuse_; }
    Vec3& getSpecular() { return specular_; }
    const Vec3& getSpecular() const { return specular_; }
    float& getShininess() { return shininess_; }
    float getShininess() const { return shininess_; }

private:
    Vec3 ambient_, diffuse_, specular_;
    float shininess_;
};

class Light {
public:
    Light() : colour_(1, 1, 1) {}
    // the direction is normalised on construction only (Light.inl:80-89), not by setDirection
    Light(const Vec3& colour, const Vec3& direction, const Vec3& position)
        : colour_(colour), direction_(direction), position_(position) { direction_.normalise(); }
    void setColour(const Vec3& v) { colour_ = v; }
    void setDirection(const Vec3& v) { direction_ = v; }
    void setPosition(const Vec3& v) { position_ = v; }
    Vec3& getColour() { return colour_; }
    const Vec3& getColour() const { return colour_; }
    Vec3& getDirection() { return direction_; }
    const Vec3& getDirection() const { return direction_; }
    Vec3& getPosition() { return position_; }
    const Vec3& getPosition() const { return position_; }

private:
    Vec3 colour_, direction_, position_;
};

} // namespace srt
