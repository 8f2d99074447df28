"""Frame exporters: PNG / OpenEXR / RGBE files written by the host codecs are decoded here by an
independent reader and compared with a numpy restatement of the reference's conversion
(image/encoding/srgb.zig:34-230, exr/exr_writer.zig:24-530, rgbe/rgbe_writer.zig:14-206)."""

import struct
import zlib

import numpy as np
import pytest

from zyg_b200 import su


def image(width=37, height=21, seed=3):
    rng = np.random.default_rng(seed)
    img = rng.random((height, width, 4), np.float32) ** 3 * 1.5
    img[0, 0] = (0, 0, 0, 0)
    img[1, 1] = (1e-4, 2.0, 0.5, 1.0)  # below the linear knee / above 1
    img[2, 2, :3] = (3e4, 1e-7, 7.0)
    return img


def linear_to_gamma(c):
    c = c.astype(np.float32)
    out = np.where(c < np.float32(0.0031308), np.float32(12.92) * c,
                   np.float32(1.055) * np.power(np.clip(c, 1e-30, None), np.float32(1.0 / 2.4), dtype=np.float32) - np.float32(0.055))
    out = np.where(c <= 0, np.float32(0), out)
    return np.where(c >= 1, np.float32(1), out).astype(np.float32)


def read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, header = 8, b"", None
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(kind + body)
        if kind == b"IHDR":
            header = struct.unpack(">IIBBBBB", body)
        if kind == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, colour, _, flt, interlace = header
    assert depth == 8 and flt == 0 and interlace == 0
    ch = {0: 1, 2: 3, 6: 4}[colour]
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * ch)
    assert not raw[:, 0].any()  # filter type None on every scanline
    return raw[:, 1:].reshape(h, w, ch)


@pytest.mark.parametrize("alpha", [False, True])
def test_png_matches_restated_conversion(tmp_path, alpha):
    img = image()
    path = str(tmp_path / "a.png")
    su.write_image(path, su.IMAGE_PNG, img, su.IMAGE_ALPHA if alpha else 0)
    got = read_png(path)
    want = (linear_to_gamma(img[..., :3]) * np.float32(255) + np.float32(0.5)).astype(np.uint8)
    # powf differs by an ulp between libm and numpy: a value that lands on .5 may round either way
    assert np.abs(got[..., :3].astype(int) - want.astype(int)).max() <= 1
    assert (got[..., :3] != want).mean() < 1e-3
    if alpha:
        a = (np.minimum(img[..., 3], 1) * np.float32(255) + np.float32(0.5)).astype(np.uint8)
        assert np.array_equal(got[..., 3], a)


def test_png_crop_and_error_diffusion(tmp_path):
    img = image(64, 16)
    crop = (5, 3, 40, 12)
    path = str(tmp_path / "c.png")
    su.write_image(path, su.IMAGE_PNG, img, su.IMAGE_ERROR_DIFFUSION, crop)
    got = read_png(path).astype(np.float64)
    inside = np.zeros(img.shape[:2], bool)
    inside[crop[1]:crop[3], crop[0]:crop[2]] = True
    assert not got[~inside].any()  # the whole frame is written, pixels outside the crop stay zero
    exact = linear_to_gamma(img[..., :3]).astype(np.float64) * 255
    # error diffusion along the row: the running sum of the quantisation error stays bounded by one level
    err = np.cumsum((exact - got)[crop[1]:crop[3], crop[0]:crop[2]], axis=1)
    assert np.abs(err).max() < 1.6


def read_exr(path):
    data = open(path, "rb").read()
    assert data[:8] == bytes([0x76, 0x2F, 0x31, 0x01, 2, 0, 0, 0])
    pos, attrs = 8, {}
    while data[pos] != 0:
        end = data.index(b"\0", pos)
        name = data[pos:end].decode()
        pos = end + 1
        end = data.index(b"\0", pos)
        kind = data[pos:end].decode()
        pos = end + 1
        (size,) = struct.unpack("<I", data[pos:pos + 4])
        attrs[name] = (kind, data[pos + 4:pos + 4 + size])
        pos += 4 + size
    pos += 1
    channels, body = [], attrs["channels"][1]
    p = 0
    while body[p] != 0:
        end = body.index(b"\0", p)
        fmt, _, xs, ys = struct.unpack("<IIII", body[end + 1:end + 17])
        channels.append((body[p:end].decode(), fmt))
        assert xs == 1 and ys == 1
        p = end + 17
    assert attrs["compression"][1] == b"\x03" and attrs["lineOrder"][1] == b"\x00"
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    display = struct.unpack("<iiii", attrs["displayWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    blocks = (h + 15) // 16
    offsets = struct.unpack(f"<{blocks}Q", data[pos:pos + 8 * blocks])
    dtype = {0: np.uint32, 1: np.float16, 2: np.float32}[channels[0][1]]
    out = np.zeros((h, w, len(channels)), np.float32)
    for b, off in enumerate(offsets):
        y, size = struct.unpack("<iI", data[off:off + 8])
        assert y == y0 + 16 * b
        rows = min(16, h - 16 * b)
        want = rows * w * len(channels) * np.dtype(dtype).itemsize
        payload = data[off + 8:off + 8 + size]
        if size < want:
            t = np.frombuffer(zlib.decompress(payload), np.uint8).astype(np.int64)
            assert t.size == want
            t[1:] = t[1:] - 128  # undo the delta predictor: t[i] = d[i] + t[i-1] - 128
            t = np.cumsum(t) & 0xFF
            half = (want + 1) // 2
            raw = np.empty(want, np.uint8)
            raw[0::2] = t[:half]
            raw[1::2] = t[half:]
        else:
            raw = np.frombuffer(payload, np.uint8)
        planes = raw.view(dtype).reshape(rows, len(channels), w)
        out[16 * b:16 * b + rows] = planes.transpose(0, 2, 1)  # (uint ids stay exact in float32 up to 2^24)
    return out, [c[0] for c in channels], (x0, y0, x1, y1), display


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("alpha", [False, True])
def test_exr_round_trip(tmp_path, half, alpha):
    img = image(45, 37)
    path = str(tmp_path / "a.exr")
    su.write_image(path, su.IMAGE_EXR, img, (su.IMAGE_HALF if half else 0) | (su.IMAGE_ALPHA if alpha else 0))
    got, names, window, display = read_exr(path)
    assert names == (["A"] if alpha else []) + ["B", "G", "R"]
    assert window == (0, 0, 44, 36) and display == (0, 0, 44, 36)
    order = ([3] if alpha else []) + [2, 1, 0]
    want = img[..., order]
    if half:
        with np.errstate(over="ignore"):
            want = want.astype(np.float16).astype(np.float32)  # round to nearest even, like @floatCast
    assert np.array_equal(got, want)


def test_exr_crop_is_the_data_window(tmp_path):
    img = image(64, 48)
    crop = (8, 5, 50, 40)
    path = str(tmp_path / "c.exr")
    su.write_image(path, su.IMAGE_EXR, img, 0, crop)
    got, _, window, display = read_exr(path)
    assert window == (8, 5, 49, 39) and display == (0, 0, 63, 47)
    assert np.array_equal(got, img[5:40, 8:50][..., [2, 1, 0]])


def read_rgbe(path):
    data = open(path, "rb").read()
    marker = b"\n\n"
    head_end = data.index(marker) + 2
    assert data.startswith(b"#?RGBE\nFORMAT=32-bit_rle_rgbe\n")
    line_end = data.index(b"\n", head_end)
    _, h, _, w = data[head_end:line_end].split()
    w, h = int(w), int(h)
    pos = line_end + 1
    out = np.zeros((h, w, 4), np.uint8)
    if w < 8 or w > 0x7FFF:
        return np.frombuffer(data[pos:pos + w * h * 4], np.uint8).reshape(h, w, 4)
    for y in range(h):
        assert data[pos] == 2 and data[pos + 1] == 2 and (data[pos + 2] << 8 | data[pos + 3]) == w
        pos += 4
        for c in range(4):
            x = 0
            while x < w:
                n = data[pos]
                if n > 128:
                    out[y, x:x + n - 128, c] = data[pos + 1]
                    x += n - 128
                    pos += 2
                else:
                    out[y, x:x + n, c] = np.frombuffer(data[pos + 1:pos + 1 + n], np.uint8)
                    x += n
                    pos += 1 + n
            assert x == w
    assert pos == len(data)
    return out


@pytest.mark.parametrize("width", [5, 40])
def test_rgbe_round_trip(tmp_path, width):
    img = image(width, 19)
    img[4:9, 3:30, :3] = 0.25  # a run for the RLE
    path = str(tmp_path / "a.hdr")
    su.write_image(path, su.IMAGE_RGBE, img)
    got = read_rgbe(path)
    rgb = np.maximum(img[..., :3], 0)
    v = rgb.max(-1)
    m, e = np.frexp(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = (m.astype(np.float32) * np.float32(256) / v).astype(np.float32)
    want = np.zeros(img.shape[:2] + (4,), np.uint8)
    ok = v >= 1e-32
    want[..., :3][ok] = (rgb[ok] * scale[ok][:, None]).astype(np.uint8)
    want[..., 3][ok] = (e[ok] + 128).astype(np.uint8)
    assert np.array_equal(got, want)
    # decoded radiance within the format's 1/128 relative precision of the brightest channel
    dec = got[..., :3].astype(np.float64) * np.exp2(got[..., 3].astype(np.float64) - 136)[..., None]
    assert np.all(np.abs(dec - rgb) <= v[..., None] / 100 + 1e-30)


def aov_image(w=40, h=28, seed=3):
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 4), np.float32)
    img[..., 0] = rng.uniform(2.0, 9.0, (h, w))      # depth / id / roughness live in the first lane
    img[..., 1:3] = rng.uniform(-1.0, 1.0, (h, w, 2))
    img[..., 3] = 1.0
    return img


def test_png_aov_encodings(tmp_path):
    """Srgb.toSrgbBuffer for the AOV encodings (image/encoding/srgb.zig:225-277): Depth and Float as one grey channel, Id as a 24-bit
    hash, Normal as 0.5 (n + 1)."""
    img = aov_image()
    unorm = lambda x: (np.clip(x, 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)  # enc.floatToUnorm8

    path = str(tmp_path / "d.png")
    img_d = img.copy()
    img_d[3, 5, 0] = np.finfo(np.float32).max  # a pixel that saw nothing: excluded from the range, black
    su.write_image(path, su.IMAGE_PNG, img_d, su.IMAGE_DEPTH)
    got = read_png(path)
    assert got.shape == (28, 40, 1)
    d = img_d[..., 0]
    lo, hi = d.min(), d[d < 2.0e9].max()
    assert np.abs(got[..., 0].astype(int) - unorm(np.float32(1.0) - (d - lo) / (hi - lo)).astype(int)).max() <= 1 and 0 == got[3, 5, 0]

    path = str(tmp_path / "f.png")
    rough = img.copy()
    rough[..., 0] = (rough[..., 0] - 2.0) / 7.0
    su.write_image(path, su.IMAGE_PNG, rough, su.IMAGE_FLOAT)
    got = read_png(path)
    assert got.shape == (28, 40, 1) and np.abs(got[..., 0].astype(int) - unorm(rough[..., 0]).astype(int)).max() <= 1

    path = str(tmp_path / "n.png")
    nrm = img.copy()
    nrm[..., 0] = np.sqrt(np.clip(1.0 - (nrm[..., 1:3] ** 2).sum(-1), 0.0, 1.0))
    su.write_image(path, su.IMAGE_PNG, nrm, su.IMAGE_NORMAL)
    got = read_png(path)
    assert got.shape == (28, 40, 3) and np.abs(got.astype(int) - unorm(np.float32(0.5) * (nrm[..., :3] + np.float32(1.0))).astype(int)).max() <= 1

    path = str(tmp_path / "i.png")
    ids = img.copy()
    ids[..., 0] = np.floor(ids[..., 0])
    su.write_image(path, su.IMAGE_PNG, ids, su.IMAGE_ID)
    got = read_png(path).astype(np.uint32)
    mid = (ids[..., 0].astype(np.uint64) * 9795927 % (1 << 32)) % 16777216
    assert np.array_equal((got[..., 0] << 16) | (got[..., 1] << 8) | got[..., 2], mid.astype(np.uint32))
    assert len(np.unique(mid)) == len(np.unique(ids[..., 0]))  # different ids, different colours


def test_exr_aov_encodings(tmp_path):
    """exr_writer.zig:42-80, 449-469: Depth = one float channel Y (never half), Id = one uint channel Y, Normal = three channels."""
    img = aov_image()
    path = str(tmp_path / "d.exr")
    su.write_image(path, su.IMAGE_EXR, img, su.IMAGE_DEPTH | su.IMAGE_HALF)
    got, names, _, _ = read_exr(path)
    assert names == ["Y"] and np.array_equal(got[..., 0], img[..., 0])
    path = str(tmp_path / "i.exr")
    su.write_image(path, su.IMAGE_EXR, img, su.IMAGE_ID)
    got, names, _, _ = read_exr(path)
    assert names == ["Y"] and np.array_equal(got[..., 0], np.floor(img[..., 0]))
    path = str(tmp_path / "n.exr")
    su.write_image(path, su.IMAGE_EXR, img, su.IMAGE_NORMAL)
    got, names, _, _ = read_exr(path)
    assert names == ["B", "G", "R"] and np.array_equal(got, img[..., [2, 1, 0]])
