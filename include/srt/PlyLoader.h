// srt/PlyLoader.h -- scene loader replacing the reference's Assimp-based loadMeshes
// (src/main-pthreads-lbuffer.cxx:518-582): PLY (binary little endian or ASCII), float x/y/z vertices,
// faces as `list uchar int|uint vertex_indices`; faces are kept in file order (Assimp keeps it for
// dragon.ply -- validated by the golden file) and non-triangles are skipped like the
// reference's `mNumIndices == 3` test. An optional 3x3 transform is applied to the vertices with the
// reference's expression order (lbuffer.cxx:562-564).
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

inline void loadMeshes(const std::string& file_name, std::vector<TriangleMesh>& mesh_set, const float* transform3x3 = nullptr)
{
    std::ifstream in(file_name, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error("Unable to open file \"" + file_name + "\"");
    std::string line, format;
    size_t n_vertices = 0, n_faces = 0;
    std::vector<std::string> vprops;
    std::string index_type = "int", count_type = "uchar";
    int section = 0;
    if (!std::getline(in, line) || line.substr(0, 3) != "ply") throw std::runtime_error(file_name + ": not a PLY file");
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string tok; ls >> tok;
        if (tok == "format") ls >> format;
        else if (tok == "element") {
            std::string what; size_t n; ls >> what >> n;
            if (what == "vertex") { n_vertices = n; section = 1; }
            else if (what == "face") { n_faces = n; section = 2; }
            else section = 3;
        } else if (tok == "property") {
            std::string t; ls >> t;
            if (section == 1) { std::string name; ls >> name; if (t != "float" && t != "float32") throw std::runtime_error(file_name + ": vertex properties must be float"); vprops.push_back(name); }
            else if (section == 2 && t == "list") { ls >> count_type >> index_type; }
        } else if (tok == "end_header") break;
    }
    if (vprops.size() < 3) throw std::runtime_error(file_name + ": vertices need x, y, z");
    const bool binary = format == "binary_little_endian";
    if (!binary && format != "ascii") throw std::runtime_error(file_name + ": unsupported PLY format " + format);

    std::vector<float> vertices(3 * n_vertices);
    std::vector<float> row(vprops.size());
    for (size_t v = 0; v < n_vertices; ++v) {
        if (binary) in.read(reinterpret_cast<char*>(row.data()), (std::streamsize)(row.size() * sizeof(float)));
        else for (float& f : row) in >> f;
        float x = row[0], y = row[1], z = row[2];
        if (transform3x3) {
            const float* m = transform3x3;
            const float tx = (m[0] * x) + (m[1] * y) + (m[2] * z), ty = (m[3] * x) + (m[4] * y) + (m[5] * z),
                        tz = (m[6] * x) + (m[7] * y) + (m[8] * z);
            x = tx; y = ty; z = tz;
        }
        vertices[3 * v] = x; vertices[3 * v + 1] = y; vertices[3 * v + 2] = z;
    }
    if (!in) throw std::runtime_error(file_name + ": truncated vertex data");
    std::vector<unsigned int> indices;
    indices.reserve(3 * n_faces);
    for (size_t f = 0; f < n_faces; ++f) {
        unsigned int cnt = 0;
        std::vector<unsigned int> idx;
        if (binary) {
            if (count_type == "uchar" || count_type == "uint8") { unsigned char c = 0; in.read(reinterpret_cast<char*>(&c), 1); cnt = c; }
            else { uint32_t c = 0; in.read(reinterpret_cast<char*>(&c), 4); cnt = c; }
            idx.resize(cnt);
            in.read(reinterpret_cast<char*>(idx.data()), (std::streamsize)(cnt * 4));      // int / uint, 4 bytes each
        } else {
            in >> cnt; idx.resize(cnt);
            for (unsigned int& i : idx) in >> i;
        }
        if (!in) throw std::runtime_error(file_name + ": truncated face data");
        if (cnt == 3) {
            for (unsigned int i : idx) if (i >= n_vertices) throw std::runtime_error(file_name + ": face index out of range");
            indices.insert(indices.end(), idx.begin(), idx.end());
        }
    }
    mesh_set.clear();
    TriangleMesh mesh;
    mesh.setGeometry(vertices, indices);
    mesh_set.push_back(mesh);
}

} // namespace srt
