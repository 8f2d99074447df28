// Frame exporters of the reference's `Image` sink (take.zig:303-331, exporting/image_sequence.zig:24-56): PNG (sRGB gamma,
// optional error diffusion), OpenEXR (ZIP, half or float, planar (A) B G R) and Radiance RGBE (.hdr). Host code: the film is
// resolved on the device, the codecs run on the resolved RGBA target like the reference's writers do.
#pragma once

#include <cstdint>
#include <vector>

namespace zyg {

// Writer.Encoding, image/image_writer.zig:17-24: what the four floats of a pixel mean and how a codec stores them. The AOV classes map
// to them through aov.Value.Class.encoding (rendering/sensor/aov/aov_value.zig:32-40).
enum class Encoding : uint32_t { Color = 0, ColorAlpha = 1, Depth = 2, Id = 3, Normal = 4, Float = 5 };

// The same with the encoding spelled out. PNG (srgb.zig:34-280): Normal = 0.5 (n + 1) as RGB, Id = a 24-bit hash of the id as RGB,
// Depth = 1 - (d - min) / (max - min) over the crop and Float = saturate(f), both as one grey channel. EXR (exr_writer.zig:42-80):
// Depth = one float channel "Y", Id = one uint channel "Y", Normal / Float = three channels like a colour.
bool encodePngAs(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], Encoding encoding,
                 bool error_diffusion);
bool encodeExrAs(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], Encoding encoding, bool half);

// image/encoding/srgb.zig:34-230 + png/png_writer.zig:33-61
bool encodePng(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], bool alpha,
               bool error_diffusion);

// image/encoding/exr/exr_writer.zig:24-164, 240-530
bool encodeExr(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4], bool alpha, bool half);

// image/encoding/rgbe/rgbe_writer.zig:14-206
bool encodeRgbe(std::vector<uint8_t>& out, const float* rgba, int32_t width, int32_t height, const int32_t crop[4]);

bool writeFile(const char* path, const std::vector<uint8_t>& bytes);

}  // namespace zyg
