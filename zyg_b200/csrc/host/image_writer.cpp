#include "image_writer.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include <zlib.h>

namespace zyg {

namespace {

// spectrum/srgb.zig linearToGamma
float linearToGamma(float c) {
    if (c <= 0.f) return 0.f;
    if (c < 0.0031308f) return 12.92f * c;
    if (c < 1.f) return 1.055f * std::pow(c, 1.f / 2.4f) - 0.055f;
    return 1.f;
}

uint8_t floatToUnorm8(float x) { return uint8_t(x * 255.f + 0.5f); }  // encoding.zig floatToUnorm8

float goldenRatio(uint32_t n) {  // srgb.zig: the per-row start of the diffused error
    const float g = float(n) * 0.618033988749894f;
    return g - std::floor(g);
}

void put32be(std::vector<uint8_t>& out, uint32_t v) {
    out.push_back(uint8_t(v >> 24));
    out.push_back(uint8_t(v >> 16));
    out.push_back(uint8_t(v >> 8));
    out.push_back(uint8_t(v));
}

void pngChunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t size) {
    put32be(out, uint32_t(size));
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data, data + size);
    put32be(out, uint32_t(crc32(0, out.data() + start, uInt(size + 4))));
}

template <typename T>
void put(std::vector<uint8_t>& out, T v) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
    out.insert(out.end(), p, p + sizeof(T));
}

void putString(std::vector<uint8_t>& out, const char* text) { out.insert(out.end(), text, text + std::strlen(text) + 1); }

// f32 -> f16, round to nearest even (Zig's @floatCast)
uint16_t floatToHalf(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t  exp  = int32_t((x >> 23) & 0xffu) - 127 + 15;
    uint32_t       man  = x & 0x7fffffu;
    if (0xffu == ((x >> 23) & 0xffu)) return uint16_t(sign | 0x7c00u | (man ? 0x200u : 0u));  // inf / nan
    if (exp >= 31) return uint16_t(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return uint16_t(sign);
        man |= 0x800000u;
        const uint32_t shift = uint32_t(14 - exp);
        uint32_t       h     = man >> shift;
        const uint32_t rem   = man & ((1u << shift) - 1u);
        const uint32_t half  = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) h += 1;
        return uint16_t(sign | h);
    }
    uint32_t h = (uint32_t(exp) << 10) | (man >> 13);
    const uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (0x1000u == rem && (h & 1u))) h += 1;  // may carry into the exponent: still correct
    return uint16_t(sign | h);
}

}  // namespace

bool writeFile(const char* path, const std::vector<uint8_t>& bytes) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    const size_t n = std::fwrite(bytes.data(), 1, bytes.size(), f);
    std::fclose(f);
    return n == bytes.size();
}


// Srgb.toSrgb + PngWriter.write, image/encoding/srgb.zig:34-230, png/png_writer.zig:33-61: the whole frame is written, pixels
// outside the crop stay zero.
bool encodePng(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], bool alpha, bool error_diffusion) {
    return encodePngAs(out, rgba, width, height, crop, alpha ? Encoding::ColorAlpha : Encoding::Color, error_diffusion);
}

bool encodePngAs(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], Encoding encoding,
                 bool error_diffusion) {
    const bool     alpha    = Encoding::ColorAlpha == encoding;
    const bool     color    = Encoding::Color == encoding || alpha;
    const uint32_t channels = alpha ? 4u : ((Encoding::Depth == encoding || Encoding::Float == encoding) ? 1u : 3u);
    std::vector<uint8_t> pixels(size_t(width) * height * channels, 0);

    float mind = FLT_MAX, maxd = 0.f;
    if (Encoding::Depth == encoding) {  // srgb.zig:50-68 (the inner loop counts from crop[1]: kept as it is)
        for (int32_t y = crop[1]; y < crop[3]; ++y) {
            size_t i = size_t(y) * width + crop[0];
            for (int32_t x = crop[1]; x < crop[2]; ++x, ++i) {
                if (i >= size_t(width) * height) break;
                const float depth = rgba[i * 4];
                mind              = std::min(mind, depth);
                if (depth < 2.14748313e+09f) maxd = std::max(maxd, depth);  // ro.RayMaxT
            }
        }
    }
    const float range = maxd - mind;
    auto        saturate = [](float x) { return std::min(std::max(x, 0.f), 1.f); };

    for (int32_t y = crop[1]; y < crop[3]; ++y) {
        float err[4];
        for (float& e : err) e = goldenRatio(uint32_t(y)) - 0.5f;
        for (int32_t x = crop[0]; x < crop[2]; ++x) {
            const float* p = rgba + (size_t(y) * width + x) * 4;
            uint8_t*     o = &pixels[(size_t(y) * width + x) * channels];
            if (!color) {
                if (Encoding::Depth == encoding) {
                    o[0] = floatToUnorm8(saturate(1.f - (p[0] - mind) / range));
                } else if (Encoding::Float == encoding) {
                    o[0] = floatToUnorm8(saturate(p[0]));
                } else if (Encoding::Id == encoding) {
                    const uint32_t id  = uint32_t(p[0]);
                    const uint32_t mid = (id * 9795927u) % 16777216u;
                    o[0] = uint8_t(mid >> 16), o[1] = uint8_t(mid >> 8), o[2] = uint8_t(mid);
                } else {  // Normal
                    for (int c = 0; c < 3; ++c) o[c] = floatToUnorm8(saturate(0.5f * (p[c] + 1.f)));
                }
                continue;
            }
            float        color4[4] = {linearToGamma(p[0]), linearToGamma(p[1]), linearToGamma(p[2]), std::min(p[3], 1.f)};
            for (uint32_t c = 0; c < channels; ++c) {
                if (error_diffusion) {
                    const float   cf = 255.f * color4[c];
                    const uint8_t ci = uint8_t(cf + err[c] + 0.5f);
                    err[c] += cf - float(ci);
                    o[c] = ci;
                } else {
                    o[c] = floatToUnorm8(color4[c]);
                }
            }
        }
    }

    // filter type 0 in front of every scanline, one zlib stream, the chunks IHDR / IDAT / IEND
    std::vector<uint8_t> raw;
    raw.reserve((size_t(width) * channels + 1) * height);
    for (int32_t y = 0; y < height; ++y) {
        raw.push_back(0);
        const uint8_t* row = &pixels[size_t(y) * width * channels];
        raw.insert(raw.end(), row, row + size_t(width) * channels);
    }
    uLongf               bound = compressBound(uLong(raw.size()));
    std::vector<uint8_t> deflated(bound);
    if (Z_OK != compress2(deflated.data(), &bound, raw.data(), uLong(raw.size()), Z_BEST_COMPRESSION)) return false;

    out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put32be(ihdr, uint32_t(width));
    put32be(ihdr, uint32_t(height));
    ihdr.push_back(8);
    ihdr.push_back(alpha ? 6 : (1 == channels ? 0 : 2));  // colour type: RGBA / grey / RGB
    ihdr.push_back(0);
    ihdr.push_back(0);
    ihdr.push_back(0);
    pngChunk(out, "IHDR", ihdr.data(), ihdr.size());
    pngChunk(out, "IDAT", deflated.data(), bound);
    pngChunk(out, "IEND", nullptr, 0);
    return true;
}

// exr_writer.zig:24-164 (header) + :240-300, 330-400, 480-530 (ZIP blocks of 16 scanlines, planar A B G R per row, byte reorder +
// delta predictor before deflate, a block that does not shrink is stored raw)
bool encodeExr(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], bool alpha, bool half) {
    return encodeExrAs(out, rgba, width, height, crop, alpha ? Encoding::ColorAlpha : Encoding::Color, half);
}

bool encodeExrAs(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], Encoding encoding, bool half_in) {
    const bool     alpha       = Encoding::ColorAlpha == encoding;
    const bool     single      = Encoding::Depth == encoding || Encoding::Id == encoding;  // exr_writer.zig:46-57
    const bool     half        = half_in && !single;
    const uint32_t channels    = alpha ? 4u : (single ? 1u : 3u);
    const uint32_t format      = Encoding::Id == encoding ? 0u : (half ? 1u : 2u);  // exr.Channel.Format: Uint 0, Half 1, Float 2
    const uint32_t scalar_size = half ? 2u : 4u;

    out = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    auto channel = [&](const char* name) {
        putString(out, name);
        put<uint32_t>(out, format);
        put<uint32_t>(out, 0);
        put<uint32_t>(out, 1);
        put<uint32_t>(out, 1);
    };
    putString(out, "channels");
    putString(out, "chlist");
    put<uint32_t>(out, channels * (2 + 4 + 4 + 4 + 4) + 1);
    if (alpha) channel("A");
    if (channels >= 3) {
        channel("B");
        channel("G");
        channel("R");
    } else {
        channel("Y");
    }
    out.push_back(0);

    putString(out, "compression");
    putString(out, "compression");
    put<uint32_t>(out, 1);
    out.push_back(3);  // exr.Compression.ZIP

    putString(out, "dataWindow");
    putString(out, "box2i");
    put<uint32_t>(out, 16);
    put<int32_t>(out, crop[0]);
    put<int32_t>(out, crop[1]);
    put<int32_t>(out, crop[2] - 1);
    put<int32_t>(out, crop[3] - 1);

    putString(out, "displayWindow");
    putString(out, "box2i");
    put<uint32_t>(out, 16);
    put<uint32_t>(out, 0);
    put<uint32_t>(out, 0);
    put<int32_t>(out, width - 1);
    put<int32_t>(out, height - 1);

    putString(out, "lineOrder");
    putString(out, "lineOrder");
    put<uint32_t>(out, 1);
    out.push_back(0);

    putString(out, "pixelAspectRatio");
    putString(out, "float");
    put<uint32_t>(out, 4);
    put<float>(out, 1.f);

    putString(out, "screenWindowCenter");
    putString(out, "v2f");
    put<uint32_t>(out, 8);
    put<float>(out, 0.f);
    put<float>(out, 0.f);

    putString(out, "screenWindowWidth");
    putString(out, "float");
    put<uint32_t>(out, 4);
    put<float>(out, 1.f);
    out.push_back(0);

    const uint32_t w = uint32_t(crop[2] - crop[0]), h = uint32_t(crop[3] - crop[1]);
    const uint32_t rows_per_block = 16;
    const uint32_t row_blocks     = (h + rows_per_block - 1) / rows_per_block;
    const uint32_t bytes_per_row  = w * channels * scalar_size;

    std::vector<std::vector<uint8_t>> blocks(row_blocks);
    std::vector<uint8_t>              block, reordered;
    for (uint32_t b = 0; b < row_blocks; ++b) {
        const uint32_t rows  = std::min(h - b * rows_per_block, rows_per_block);
        const uint32_t bytes = rows * bytes_per_row;
        block.assign(bytes, 0);
        for (uint32_t row = 0; row < rows; ++row) {
            const uint32_t y = uint32_t(crop[1]) + b * rows_per_block + row;
            for (uint32_t x = 0; x < w; ++x) {
                const float* p = rgba + (size_t(y) * width + uint32_t(crop[0]) + x) * 4;
                for (uint32_t c = 0; c < channels; ++c) {
                    const float  v = p[channels - 1 - c];  // planes in the order (A) B G R, or the one plane Y
                    const size_t o = (size_t(row) * w * channels + size_t(w) * c + x) * scalar_size;
                    if (0u == format) {  // blockUint, exr_writer.zig:449-469
                        const uint32_t uv = uint32_t(v);
                        std::memcpy(&block[o], &uv, 4);
                    } else if (half) {
                        const uint16_t hv = floatToHalf(v);
                        std::memcpy(&block[o], &hv, 2);
                    } else {
                        std::memcpy(&block[o], &v, 4);
                    }
                }
            }
        }
        // reorder: even bytes first, odd bytes from the middle; then the delta predictor
        reordered.assign(bytes, 0);
        size_t t1 = 0, t2 = (size_t(bytes) + 1) / 2;
        for (size_t cur = 0; cur < bytes;) {
            reordered[t1++] = block[cur++];
            if (cur < bytes) reordered[t2++] = block[cur++];
        }
        uint32_t prev = reordered[0];
        for (size_t t = 1; t < bytes; ++t) {
            const uint32_t cur = reordered[t];
            reordered[t]       = uint8_t(cur - prev + (128 + 256));
            prev               = cur;
        }
        uLongf               size = compressBound(bytes);
        std::vector<uint8_t> deflated(size);
        if (Z_OK != compress2(deflated.data(), &size, reordered.data(), bytes, Z_BEST_COMPRESSION)) return false;
        if (size >= bytes) {
            blocks[b] = block;  // stored raw
        } else {
            deflated.resize(size);
            blocks[b] = std::move(deflated);
        }
    }

    uint64_t offset = out.size() + uint64_t(row_blocks) * 8;
    for (uint32_t b = 0; b < row_blocks; ++b) {
        put<uint64_t>(out, offset);
        offset += 4 + 4 + blocks[b].size();
    }
    for (uint32_t b = 0; b < row_blocks; ++b) {
        put<uint32_t>(out, uint32_t(crop[1]) + b * rows_per_block);
        put<uint32_t>(out, uint32_t(blocks[b].size()));
        out.insert(out.end(), blocks[b].begin(), blocks[b].end());
    }
    return true;
}


namespace {

// rgbe_writer.zig:183-205
void floatToRgbe(const float* c, uint8_t rgbe[4]) {
    float v = c[0];
    if (c[1] > v) v = c[1];
    if (c[2] > v) v = c[2];
    if (v < 1.0e-32f) {
        rgbe[0] = rgbe[1] = rgbe[2] = rgbe[3] = 0;
        return;
    }
    int         exponent;
    const float significand = std::frexp(v, &exponent);
    v       = significand * 256.f / v;
    rgbe[0] = uint8_t(c[0] * v);
    rgbe[1] = uint8_t(c[1] * v);
    rgbe[2] = uint8_t(c[2] * v);
    rgbe[3] = uint8_t(exponent + 128);
}

// rgbe_writer.zig:119-181: runs of at least 4 equal bytes become (128 + n, value), everything else literal chunks of up to 128
void rgbeRle(std::vector<uint8_t>& out, const uint8_t* data, uint32_t len) {
    constexpr uint32_t kMinRun = 4;
    uint32_t           current = 0;
    while (current < len) {
        uint32_t begin_run = current, run_count = 0, old_run_count = 0;
        while (run_count < kMinRun && begin_run < len) {
            begin_run += run_count;
            old_run_count = run_count;
            run_count     = 1;
            while (begin_run + run_count < len && run_count < 127 && data[begin_run] == data[begin_run + run_count]) ++run_count;
        }
        if (old_run_count > 1 && old_run_count == begin_run - current) {
            out.push_back(uint8_t(128 + old_run_count));
            out.push_back(data[current]);
            current = begin_run;
        }
        while (current < begin_run) {
            const uint32_t n = std::min(begin_run - current, 128u);
            out.push_back(uint8_t(n));
            out.insert(out.end(), data + current, data + current + n);
            current += n;
        }
        if (run_count >= kMinRun) {
            out.push_back(uint8_t(128 + run_count));
            out.push_back(data[begin_run]);
            current += run_count;
        }
    }
}

}  // namespace

bool encodeRgbe(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4]) {
    out.clear();
    char header[96];
    const int n = std::snprintf(header, sizeof(header), "#?RGBE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", height, width);
    out.insert(out.end(), header, header + n);

    auto clamped = [&](size_t i, float c[3]) {
        for (int k = 0; k < 3; ++k) c[k] = 0.f < rgba[i * 4 + k] ? rgba[i * 4 + k] : 0.f;  // math.max4(p, 0)
    };

    if (width < 8 || width > 0x7fff) {  // run-length encoding is not allowed: flat RGBE quadruples (:94-117)
        for (int32_t y = 0; y < height; ++y) {
            for (int32_t x = 0; x < width; ++x) {
                uint8_t rgbe[4] = {0, 0, 0, 0};
                if (!(y < crop[1] || y >= crop[3] || x < crop[0] || x >= crop[2])) {
                    float c[3];
                    clamped(size_t(y) * width + x, c);
                    floatToRgbe(c, rgbe);
                }
                out.insert(out.end(), rgbe, rgbe + 4);
            }
        }
        return true;
    }

    const uint32_t       w = uint32_t(width);
    std::vector<uint8_t> row(size_t(w) * 4);
    for (int32_t y = 0; y < height; ++y) {
        const uint8_t info[4] = {2, 2, uint8_t(w >> 8), uint8_t(w & 0xff)};
        out.insert(out.end(), info, info + 4);
        if (y < crop[1] || y >= crop[3]) {
            std::fill(row.begin(), row.end(), uint8_t(0));
        } else {
            for (uint32_t x = 0; x < w; ++x) {
                if (int32_t(x) < crop[0] || int32_t(x) >= crop[2]) {
                    // the reference zeroes the interleaved position here (:66-70), not the planar one: kept as it is
                    row[x * 4 + 0] = row[x * 4 + 1] = row[x * 4 + 2] = row[x * 4 + 3] = 0;
                } else {
                    float   c[3];
                    uint8_t rgbe[4];
                    clamped(size_t(y) * w + x, c);
                    floatToRgbe(c, rgbe);
                    for (uint32_t k = 0; k < 4; ++k) row[x + w * k] = rgbe[k];
                }
            }
        }
        for (uint32_t k = 0; k < 4; ++k) rgbeRle(out, row.data() + size_t(k) * w, w);
    }
    return true;
}

}  // namespace zyg
