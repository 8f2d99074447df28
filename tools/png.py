"""Tiny PNG writer (zlib + struct) for eyeballing renders; no imaging dependency."""
import struct
import zlib

import numpy as np


def write_png(path, rgb):
    """rgb: float array (H, W, 3+) in linear sRGB; applies the sRGB transfer curve and writes 8-bit RGB."""
    c = np.clip(np.asarray(rgb, np.float64)[..., :3], 0.0, 1.0)
    c = np.where(c < 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055)
    img = (c * 255.0 + 0.5).astype(np.uint8)
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
