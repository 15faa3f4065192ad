// srt/PlyLoader.h -- scene loader replacing the reference's Assimp-based loadMeshes
// (src/main-pthreads-lbuffer.cxx:518-582).  Reads PLY files: ascii, binary_little_endian and binary_big_endian; any
// element / property layout (scalar properties of every PLY type, list properties) -- the vertex element supplies
// x, y, z (other properties, e.g. normals, colours, confidence, are read past), the face element its
// vertex_indices / vertex_index list (other properties are read past), every other element is skipped.
// Faces are kept in file order (Assimp keeps it for dragon.ply -- validated by the golden file)
// and non-triangles are skipped like the reference's `mNumIndices == 3` test.  float coordinates are taken bit for
// bit; double ones are narrowed.  An optional 3x3 transform is applied to the vertices with the reference's
// expression order (lbuffer.cxx:562-564).
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "TriangleMesh.h"

namespace srt {
namespace ply_detail {

enum Type { kNone, kI8, kU8, kI16, kU16, kI32, kU32, kF32, kF64 };

inline Type parse_type(const std::string& t)
{
    if (t == "char" || t == "int8") return kI8;
    if (t == "uchar" || t == "uint8") return kU8;
    if (t == "short" || t == "int16") return kI16;
    if (t == "ushort" || t == "uint16") return kU16;
    if (t == "int" || t == "int32") return kI32;
    if (t == "uint" || t == "uint32") return kU32;
    if (t == "float" || t == "float32") return kF32;
    if (t == "double" || t == "float64") return kF64;
    return kNone;
}

inline size_t size_of(Type t) { return t == kF64 ? 8 : (t == kI32 || t == kU32 || t == kF32) ? 4 : (t == kI16 || t == kU16) ? 2 : 1; }

struct Property { std::string name; bool is_list = false; Type type = kNone, count_type = kNone; };
struct Element { std::string name; size_t count = 0; std::vector<Property> props; };

// One value of the data section.  float32 values are returned through `as_float` bit for bit (a detour through double
// would be exact as well, but NaN payloads need not survive it).
class Reader {
public:
    Reader(std::istream& in, bool binary, bool big_endian) : in_(in), binary_(binary), swap_(big_endian != host_big_endian()) {}

    double number(Type t, float* as_float = nullptr)
    {
        if (!binary_) {
            double v = 0.0;
            if (t == kF32 && as_float) { in_ >> *as_float; return *as_float; }
            in_ >> v;
            if (as_float) *as_float = (float)v;
            return v;
        }
        unsigned char b[8] = { 0 };
        const size_t n = size_of(t);
        in_.read(reinterpret_cast<char*>(b), (std::streamsize)n);
        if (swap_) for (size_t i = 0; i < n / 2; ++i) std::swap(b[i], b[n - 1 - i]);
        double v = 0.0;
        switch (t) {
            case kI8: { int8_t x; std::memcpy(&x, b, 1); v = x; break; }
            case kU8: { uint8_t x; std::memcpy(&x, b, 1); v = x; break; }
            case kI16: { int16_t x; std::memcpy(&x, b, 2); v = x; break; }
            case kU16: { uint16_t x; std::memcpy(&x, b, 2); v = x; break; }
            case kI32: { int32_t x; std::memcpy(&x, b, 4); v = x; break; }
            case kU32: { uint32_t x; std::memcpy(&x, b, 4); v = x; break; }
            case kF32: { float x; std::memcpy(&x, b, 4); if (as_float) { *as_float = x; return x; } v = x; break; }
            case kF64: { double x; std::memcpy(&x, b, 8); v = x; break; }
            default: break;
        }
        if (as_float) *as_float = (float)v;
        return v;
    }

    bool ok() const { return !in_.fail(); }

private:
    static bool host_big_endian() { const uint16_t one = 1; unsigned char c; std::memcpy(&c, &one, 1); return c == 0; }
    std::istream& in_;
    bool binary_, swap_;
};

} // namespace ply_detail

inline void loadMeshes(const std::string& file_name, std::vector<TriangleMesh>& mesh_set, const float* transform3x3 = nullptr)
{
    using namespace ply_detail;
    std::ifstream in(file_name, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error("Unable to open file \"" + file_name + "\"");
    std::string line, format;
    std::vector<Element> elements;
    bool header_done = false;
    if (!std::getline(in, line) || line.substr(0, 3) != "ply") throw std::runtime_error(file_name + ": not a PLY file");
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string tok; ls >> tok;
        if (tok == "format") ls >> format;
        else if (tok == "element") {
            Element e; ls >> e.name >> e.count;
            if (ls.fail()) throw std::runtime_error(file_name + ": malformed element line");
            elements.push_back(e);
        } else if (tok == "property") {
            if (elements.empty()) throw std::runtime_error(file_name + ": property before any element");
            Property p; std::string t; ls >> t;
            if (t == "list") {
                std::string ct, it; ls >> ct >> it >> p.name;
                p.is_list = true; p.count_type = parse_type(ct); p.type = parse_type(it);
                if (p.count_type == kNone || p.count_type == kF32 || p.count_type == kF64)
                    throw std::runtime_error(file_name + ": unsupported list count type " + ct);
            } else {
                p.type = parse_type(t); ls >> p.name;
            }
            if (p.type == kNone || ls.fail()) throw std::runtime_error(file_name + ": unsupported property \"" + line + "\"");
            elements.back().props.push_back(p);
        } else if (tok == "end_header") { header_done = true; break; }
    }
    if (!header_done) throw std::runtime_error(file_name + ": no end_header");
    const bool binary = format == "binary_little_endian" || format == "binary_big_endian";
    if (!binary && format != "ascii") throw std::runtime_error(file_name + ": unsupported PLY format " + format);
    Reader rd(in, binary, format == "binary_big_endian");

    std::vector<float> vertices;
    std::vector<unsigned int> indices;
    size_t n_vertices = 0;
    bool have_vertices = false, have_faces = false;
    for (const Element& e : elements) {
        if (e.name == "vertex") {
            // x, y, z by name; a file that does not name them supplies them as its first three scalar properties
            int ix = -1, iy = -1, iz = -1, scalars = 0;
            for (size_t k = 0; k < e.props.size(); ++k) {
                if (e.props[k].is_list) continue;
                ++scalars;
                if (e.props[k].name == "x") ix = (int)k; else if (e.props[k].name == "y") iy = (int)k; else if (e.props[k].name == "z") iz = (int)k;
            }
            if (scalars < 3) throw std::runtime_error(file_name + ": vertices need x, y, z");
            if (ix < 0 || iy < 0 || iz < 0) {
                int found = 0;
                for (size_t k = 0; k < e.props.size() && found < 3; ++k)
                    if (!e.props[k].is_list) { (found == 0 ? ix : found == 1 ? iy : iz) = (int)k; ++found; }
            }
            n_vertices = e.count; have_vertices = true;
            vertices.resize(3 * n_vertices);
            for (size_t v = 0; v < n_vertices; ++v) {
                float x = 0, y = 0, z = 0;
                for (size_t k = 0; k < e.props.size(); ++k) {
                    const Property& p = e.props[k];
                    if (p.is_list) { const size_t n = (size_t)rd.number(p.count_type); for (size_t i = 0; i < n && rd.ok(); ++i) rd.number(p.type); continue; }
                    float f = 0;
                    rd.number(p.type, &f);
                    if ((int)k == ix) x = f; else if ((int)k == iy) y = f; else if ((int)k == iz) z = f;
                }
                if (transform3x3) {
                    const float* m = transform3x3;
                    const float tx = (m[0] * x) + (m[1] * y) + (m[2] * z), ty = (m[3] * x) + (m[4] * y) + (m[5] * z),
                                tz = (m[6] * x) + (m[7] * y) + (m[8] * z);
                    x = tx; y = ty; z = tz;
                }
                vertices[3 * v] = x; vertices[3 * v + 1] = y; vertices[3 * v + 2] = z;
                if (!rd.ok()) throw std::runtime_error(file_name + ": truncated vertex data");
            }
        } else if (e.name == "face") {
            if (!have_vertices) throw std::runtime_error(file_name + ": faces before vertices");
            int il = -1;
            for (size_t k = 0; k < e.props.size(); ++k)
                if (e.props[k].is_list && (e.props[k].name == "vertex_indices" || e.props[k].name == "vertex_index")) il = (int)k;
            if (il < 0)
                for (size_t k = 0; k < e.props.size() && il < 0; ++k) if (e.props[k].is_list) il = (int)k;
            if (il < 0 && e.count) throw std::runtime_error(file_name + ": faces have no index list");
            have_faces = true;
            indices.reserve(3 * e.count);
            std::vector<unsigned int> idx;
            for (size_t f = 0; f < e.count; ++f) {
                for (size_t k = 0; k < e.props.size(); ++k) {
                    const Property& p = e.props[k];
                    if (!p.is_list) { rd.number(p.type); continue; }
                    const double cnt = rd.number(p.count_type);
                    if (!rd.ok() || cnt < 0 || cnt > 1e6) throw std::runtime_error(file_name + ": truncated face data");
                    const size_t n = (size_t)cnt;
                    if ((int)k != il) { for (size_t i = 0; i < n && rd.ok(); ++i) rd.number(p.type); continue; }
                    idx.resize(n);
                    for (size_t i = 0; i < n; ++i) {
                        const double v = rd.number(p.type);
                        if (rd.ok() && (v < 0 || v >= (double)n_vertices)) throw std::runtime_error(file_name + ": face index out of range");
                        idx[i] = (unsigned int)v;
                    }
                    if (!rd.ok()) throw std::runtime_error(file_name + ": truncated face data");
                    if (n == 3) indices.insert(indices.end(), idx.begin(), idx.end());
                }
                if (!rd.ok()) throw std::runtime_error(file_name + ": truncated face data");
            }
        } else {
            for (size_t r = 0; r < e.count; ++r)
                for (const Property& p : e.props) {
                    if (!p.is_list) { rd.number(p.type); continue; }
                    const size_t n = (size_t)rd.number(p.count_type);
                    for (size_t i = 0; i < n && rd.ok(); ++i) rd.number(p.type);
                }
            if (!rd.ok()) throw std::runtime_error(file_name + ": truncated data in element " + e.name);
        }
    }
    if (!have_vertices) throw std::runtime_error(file_name + ": vertices need x, y, z");
    (void)have_faces;
    mesh_set.clear();
    TriangleMesh mesh;
    mesh.setGeometry(vertices, indices);
    mesh_set.push_back(mesh);
}

} // namespace srt
