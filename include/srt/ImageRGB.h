// srt/ImageRGB.h -- 8-bit RGB image with the API of the reference's older Image class
// (include/oldimage.h:35-236, include/oldimage.inl, src/Image.bakk.cxx), which the shaded drivers
// (main-omp / main-mpi / main-pthreads-rows) render into and which TriangleMesh uses as a texture.
// libjpeg is not available here, so files are read/written as binary PPM (P6) or TGA instead of JPEG.
#pragma once
#include "../srt.h"
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace srt {

class ImageRGB {
public:
    ImageRGB() = default;
    ImageRGB(unsigned int width, unsigned int height, unsigned char r = 0, unsigned char g = 0, unsigned char b = 0) { setSize(width, height, r, g, b); }
    explicit ImageRGB(const std::string& ppm_file) { loadPPMFile(ppm_file); }

    unsigned int getWidth() const { return width_; }
    unsigned int getHeight() const { return height_; }
    void getSize(unsigned int& w, unsigned int& h) const { w = width_; h = height_; }
    unsigned char* getData() const { return const_cast<unsigned char*>(data_.data()); }
    void destroy() { data_.clear(); width_ = height_ = 0; }

    // index = j * width * 3 + i * 3 (oldimage.inl:164-217); out of range throws std::out_of_range
    void setPixel(unsigned int i, unsigned int j, unsigned char r, unsigned char g, unsigned char b)
    { unsigned char* p = &data_[index(i, j)]; p[0] = r; p[1] = g; p[2] = b; }
    void getPixel(unsigned int i, unsigned int j, unsigned char& r, unsigned char& g, unsigned char& b) const
    { const unsigned char* p = &data_[index(i, j)]; r = p[0]; g = p[1]; b = p[2]; }

    void setSize(unsigned int width, unsigned int height, unsigned char r = 0, unsigned char g = 0, unsigned char b = 0)
    {
        width_ = width; height_ = height;
        data_.resize((size_t)width * height * 3);
        for (size_t k = 0; k < data_.size(); k += 3) { data_[k] = r; data_[k + 1] = g; data_[k + 2] = b; }
    }

    void savePPMFile(const std::string& name) const
    {
        std::ofstream f(name, std::ios::binary);
        if (!f.is_open()) throw std::runtime_error("Cannot create the file " + name);
        f << "P6\n" << width_ << " " << height_ << "\n255\n";
        f.write(reinterpret_cast<const char*>(data_.data()), (std::streamsize)data_.size());
    }
    void loadPPMFile(const std::string& name)
    {
        std::ifstream f(name, std::ios::binary);
        std::string magic; unsigned int w = 0, h = 0, maxv = 0;
        if (!f.is_open() || !(f >> magic >> w >> h >> maxv) || magic != "P6" || maxv != 255)
            throw std::runtime_error("Cannot read the PPM file " + name);
        f.get();
        setSize(w, h);
        f.read(reinterpret_cast<char*>(data_.data()), (std::streamsize)data_.size());
        if (!f) throw std::runtime_error("Truncated PPM file " + name);
    }
    // uncompressed 24-bit TGA, header layout of src/Image.bakk.cxx:230-290
    // old Image::saveJPEGFile (src/Image.bakk.cxx:170-226): quality 100 -- encoded by nvJPEG on the GPU of `ctx`
    void saveJPEGFile(srt_ctx* ctx, const std::string& name) const
    {
        if (srt_write_jpeg(ctx, name.c_str(), getData(), getWidth(), getHeight(), 100) != SRT_OK)
            throw std::runtime_error(std::string("saveJPEGFile: ") + srt_last_error(ctx));
    }

    void saveTGAFile(const std::string& name) const
    {
        std::FILE* f = std::fopen(name.c_str(), "wb");
        if (!f) throw std::runtime_error("Cannot create the file " + name);
        const unsigned char header[18] = { 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                           (unsigned char)(width_ & 0xFF), (unsigned char)((width_ >> 8) & 0xFF),
                                           (unsigned char)(height_ & 0xFF), (unsigned char)((height_ >> 8) & 0xFF), 24, 0 };
        std::fwrite(header, 1, sizeof header, f);
        for (unsigned int r = height_; r-- > 0;)
            for (unsigned int c = 0; c < width_; ++c) {
                const unsigned char* p = &data_[((size_t)r * width_ + c) * 3];
                const unsigned char bgr[3] = { p[2], p[1], p[0] };
                std::fwrite(bgr, 1, 3, f);
            }
        std::fclose(f);
    }

private:
    size_t index(unsigned int i, unsigned int j) const
    {
        if (i < width_ && j < height_) return ((size_t)j * width_ + i) * 3;
        std::stringstream msg;
        msg << "pixel (" << i << ", " << j << ") outside a " << width_ << " x " << height_ << " image";
        throw std::out_of_range(msg.str());
    }
    std::vector<unsigned char> data_;
    unsigned int width_ = 0, height_ = 0;
};

} // namespace srt
