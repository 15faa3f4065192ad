// srt/Image.h -- float greyscale image, the reference's current Image API
// (include/Image.h:35-215, include/Image.inl, src/Image.cxx).
//
// Kept: row-major float storage (index j*width + i), bounds-checked setPixel/getPixel throwing
// std::out_of_range (Image.inl:139-185), deep copies, saveTextFile byte-compatible with the
// reference's golden file (tab separated, default ostream float formatting, newline between rows,
// none after the last; src/Image.cxx:210-235).
// NOT copied: the reference's applyLUT indexes a 1-channel buffer as if it had 3 channels and
// overwrites one output byte three times (Image.inl:189-216), so its JPEG/TGA output is wrong.
// Here applyLUT is the intended mapping round(255 (v - vmin) / (vmax - vmin)) clamped to [0, 255],
// replicated to RGB; saveTGAFile writes that as an uncompressed 24-bit TGA with the reference's
// header layout (src/Image.cxx:148-206); savePGMFile is an extra 8-bit writer.
#pragma once
#include "../srt.h"
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace srt {

class Image {
public:
    Image() = default;
    Image(unsigned int width, unsigned int height, float value = 0.0f) { setSize(width, height, value); }
    Image(const Image&) = default;
    Image& operator=(const Image&) = default;

    void destroy() { data_.clear(); data_.shrink_to_fit(); width_ = height_ = 0; }

    unsigned int getWidth() const { return width_; }
    unsigned int getHeight() const { return height_; }
    void getSize(unsigned int& w, unsigned int& h) const { w = width_; h = height_; }
    float* getData() const { return const_cast<float*>(data_.data()); }

    void setPixel(unsigned int i, unsigned int j, float value) { data_[index(i, j)] = value; }
    void getPixel(unsigned int i, unsigned int j, float& value) const { value = data_[index(i, j)]; }

    void saveTextFile(const std::string& name) const
    {
        std::ofstream f(name);
        if (!f.is_open()) throw std::runtime_error("Cannot create the file " + name);
        for (unsigned int r = 0; r < height_; ++r) {
            for (unsigned int c = 0; c < width_; ++c) {
                f << data_[(size_t)r * width_ + c];
                if (c + 1 < width_) f << "\t";
            }
            if (r + 1 < height_) f << std::endl;
        }
    }

    // 3 bytes (R = G = B) per pixel
    std::vector<unsigned char> applyLUT(float vmin, float vmax) const
    {
        std::vector<unsigned char> out(3 * data_.size());
        for (size_t k = 0; k < data_.size(); ++k) {
            const float v = data_[k];
            unsigned char g;
            if (!(v >= vmin)) g = 0;                       // also NaN
            else if (v > vmax) g = 255;
            else g = (unsigned char)std::lround(255.0 * (v - vmin) / (vmax - vmin));
            out[3 * k] = out[3 * k + 1] = out[3 * k + 2] = g;
        }
        return out;
    }

    void saveTGAFile(const std::string& name, float vmin = 0.0f, float vmax = 1.0f) const
    {
        std::FILE* f = std::fopen(name.c_str(), "wb");
        if (!f) throw std::runtime_error("Cannot create the file " + name);
        const unsigned char header[18] = { 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                           (unsigned char)(width_ & 0xFF), (unsigned char)((width_ >> 8) & 0xFF),
                                           (unsigned char)(height_ & 0xFF), (unsigned char)((height_ >> 8) & 0xFF), 24, 0 };
        std::fwrite(header, 1, sizeof header, f);
        const std::vector<unsigned char> rgb = applyLUT(vmin, vmax);
        for (unsigned int r = height_; r-- > 0;)             // TGA rows are stored bottom-up
            for (unsigned int c = 0; c < width_; ++c) {
                const unsigned char* p = &rgb[3 * ((size_t)r * width_ + c)];
                const unsigned char bgr[3] = { p[2], p[1], p[0] };
                std::fwrite(bgr, 1, 3, f);
            }
        std::fclose(f);
    }

    // Image::saveJPEGFile (src/Image.cxx:85-146): LUT to 8 bits, 3 components, quality 100 -- encoded by nvJPEG on the
    // GPU of `ctx` (srt_write_jpeg); the LUT is the intended one (see applyLUT above)
    void saveJPEGFile(srt_ctx* ctx, const std::string& name, float vmin = 0.0f, float vmax = 1.0f) const
    {
        const std::vector<unsigned char> rgb = applyLUT(vmin, vmax);
        if (srt_write_jpeg(ctx, name.c_str(), rgb.data(), width_, height_, 100) != SRT_OK)
            throw std::runtime_error(std::string("saveJPEGFile: ") + srt_last_error(ctx));
    }

    void savePGMFile(const std::string& name, float vmin = 0.0f, float vmax = 1.0f) const
    {
        std::ofstream f(name, std::ios::binary);
        if (!f.is_open()) throw std::runtime_error("Cannot create the file " + name);
        f << "P5\n" << width_ << " " << height_ << "\n255\n";
        const std::vector<unsigned char> rgb = applyLUT(vmin, vmax);
        for (size_t k = 0; k < data_.size(); ++k) f.put((char)rgb[3 * k]);
    }

    // (re)allocate; like the reference, nothing happens when the size is unchanged (src/Image.cxx:239-263)
    void setSize(unsigned int width, unsigned int height, float value = 0.0f)
    {
        if (width == width_ && height == height_) return;
        width_ = width; height_ = height;
        data_.assign((size_t)width * height, value);
    }

private:
    size_t index(unsigned int i, unsigned int j) const
    {
        if (i < width_ && j < height_) return (size_t)j * width_ + i;
        std::stringstream msg;
        msg << "pixel (" << i << ", " << j << ") outside a " << width_ << " x " << height_ << " image";
        throw std::out_of_range(msg.str());
    }

    std::vector<float> data_;
    unsigned int width_ = 0, height_ = 0;
};

} // namespace srt
