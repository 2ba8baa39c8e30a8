"""Image output for human-checkable artefacts: PNG (8-bit, what apps/SimpleViewer/ReferenceImages
holds) and OpenEXR (half or float, uncompressed scan lines) writers plus a PNG reader, numpy + zlib only. The reference writes
its screenshots through stb_image_write / tinyexr (extensions/StbImageLoader, extensions/TinyExr); these produce the same file
formats from the arrays bpt_resolve_tonemapped / bpt_resolve_float4 return.
"""
import struct
import zlib

import numpy as np


def _chunk(tag, payload):
    return struct.pack(">I", len(payload)) + tag + payload + struct.pack(">I", zlib.crc32(tag + payload) & 0xFFFFFFFF)


def write_png(path, pixels, flip_vertically=True):
    """pixels: (H, W, 3 or 4) uint8. Row 0 of the renderer's frame is the BOTTOM row (OpenGL / OptiX convention), PNG stores
    the top row first: flipped by default."""
    p = np.ascontiguousarray(pixels)
    assert p.dtype == np.uint8 and p.ndim == 3 and p.shape[2] in (3, 4)
    if flip_vertically:
        p = p[::-1]
    h, w, c = p.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), p.reshape(h, w * c)], axis=1).tobytes()  # filter type 0 on every row
    header = struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 6, 0, 0, 0)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", header) + _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b""))


def read_png(path, flip_vertically=True):
    """8-bit RGB / RGBA / grey PNGs without interlacing -> (H, W, C) uint8 (bottom row first by default)."""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG file"
    pos, idat, header = 8, b"", None
    while pos < len(data):
        length, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + length]
        if tag == b"IHDR":
            header = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + length
    w, h, depth, colour, _, _, interlace = header
    assert depth == 8 and interlace == 0 and colour in (0, 2, 4, 6), "only 8-bit non-interlaced grey / RGB(A) PNGs"
    c = {0: 1, 2: 3, 4: 2, 6: 4}[colour]
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * c)
    out = np.zeros((h, w * c), np.uint8)
    previous = np.zeros(w * c, np.int32)
    for y in range(h):
        kind, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if kind == 0:
            current = line
        elif kind == 2:
            current = (line + previous) & 255
        else:  # 1 (sub), 3 (average), 4 (Paeth): left neighbours make the row sequential per pixel, vectorised over channels
            current = np.zeros(w * c, np.int32)
            for x in range(w):
                s = slice(x * c, (x + 1) * c)
                left = current[(x - 1) * c:x * c] if x else np.zeros(c, np.int32)
                up = previous[s]
                up_left = previous[(x - 1) * c:x * c] if x else np.zeros(c, np.int32)
                if kind == 1:
                    predictor = left
                elif kind == 3:
                    predictor = (left + up) >> 1
                else:
                    estimate = left + up - up_left
                    pa, pb, pc = np.abs(estimate - left), np.abs(estimate - up), np.abs(estimate - up_left)
                    predictor = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, up_left))
                current[s] = (line[s] + predictor) & 255
        out[y] = current
        previous = current
    out = out.reshape(h, w, c)
    return out[::-1].copy() if flip_vertically else out


def write_exr(path, pixels, half=True, flip_vertically=True):
    """pixels: (H, W, 3 or 4) float. Single-part scan-line OpenEXR 2.0, no compression, channels A B G R in that (alphabetical)
    order, HALF or FLOAT samples."""
    p = np.asarray(pixels, np.float32)
    assert p.ndim == 3 and p.shape[2] in (3, 4)
    if flip_vertically:
        p = p[::-1]
    h, w, c = p.shape
    names = ["B", "G", "R"] if c == 3 else ["A", "B", "G", "R"]
    source = {"R": 0, "G": 1, "B": 2, "A": 3}
    pixel_type, dtype, size = (1, np.float16, 2) if half else (2, np.float32, 4)

    def attribute(name, kind, payload):
        return name.encode() + b"\0" + kind.encode() + b"\0" + struct.pack("<i", len(payload)) + payload

    channels = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", pixel_type, 0, 0, 0, 0, 1, 1) for n in names) + b"\0"
    window = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = (struct.pack("<II", 20000630, 2) + attribute("channels", "chlist", channels) + attribute("compression", "compression", b"\0")
              + attribute("dataWindow", "box2i", window) + attribute("displayWindow", "box2i", window) + attribute("lineOrder", "lineOrder", b"\0")
              + attribute("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attribute("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
              + attribute("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    line_bytes = len(names) * w * size
    table_start = len(header) + 8 * h
    offsets = struct.pack(f"<{h}Q", *[table_start + y * (8 + line_bytes) for y in range(h)])
    with open(path, "wb") as f:
        f.write(header + offsets)
        for y in range(h):
            planes = b"".join(np.ascontiguousarray(p[y, :, source[n]]).astype(dtype).tobytes() for n in names)
            f.write(struct.pack("<ii", y, line_bytes) + planes)


def read_exr(path, flip_vertically=True):
    """Reads back what write_exr wrote (uncompressed scan lines, HALF or FLOAT): -> (H, W, C) float32, channels R G B [A]."""
    data = open(path, "rb").read()
    magic, version = struct.unpack("<II", data[:8])
    assert magic == 20000630 and (version & 0xFF) == 2
    pos, attributes = 8, {}
    while data[pos] != 0:
        end = data.index(b"\0", pos); name = data[pos:end].decode(); pos = end + 1
        end = data.index(b"\0", pos); pos = end + 1
        (length,) = struct.unpack("<i", data[pos:pos + 4]); pos += 4
        attributes[name] = data[pos:pos + length]; pos += length
    pos += 1
    assert attributes["compression"] == b"\0", "only uncompressed files"
    x0, y0, x1, y1 = struct.unpack("<iiii", attributes["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    names, types, cp, chl = [], [], 0, attributes["channels"]
    while chl[cp] != 0:
        end = chl.index(b"\0", cp); names.append(chl[cp:end].decode()); cp = end + 1
        types.append(struct.unpack("<i", chl[cp:cp + 4])[0]); cp += 16
    offsets = struct.unpack(f"<{h}Q", data[pos:pos + 8 * h])
    out = np.zeros((h, w, len(names)), np.float32)
    order = {"R": 0, "G": 1, "B": 2, "A": 3}
    for y in range(h):
        p = offsets[y] + 8
        for n, t in zip(names, types):
            dtype, size = (np.float16, 2) if t == 1 else (np.float32, 4)
            out[y, :, order[n]] = np.frombuffer(data[p:p + w * size], dtype).astype(np.float32); p += w * size
    return out[::-1].copy() if flip_vertically else out
