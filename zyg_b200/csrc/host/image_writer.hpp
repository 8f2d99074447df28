// Frame exporters of the reference's `Image` sink (take.zig:303-331, exporting/image_sequence.zig:24-56): PNG (sRGB gamma,
// optional error diffusion), OpenEXR (ZIP, half or float, planar (A) B G R) and Radiance RGBE (.hdr). Host code: the film is
// resolved on the device, the codecs run on the resolved RGBA target like the reference's writers do.
#pragma once

#include <cstdint>
#include <vector>

namespace zyg {

// image/encoding/srgb.zig:34-230 + png/png_writer.zig:33-61
bool encodePng(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], bool alpha,
               bool error_diffusion);

// image/encoding/exr/exr_writer.zig:24-164, 240-530
bool encodeExr(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], bool alpha, bool half);

// image/encoding/rgbe/rgbe_writer.zig:14-206
bool encodeRgbe(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4]);

bool writeFile(const char* path, const std::vector<uint8_t>& bytes);

}  // namespace zyg
