"""The .klg container as kintinuous_b200/klg.py writes it for the GPU tests, against the layout RawLogReader parses
(src/utils/RawLogReader.cpp:29, :54-66): int32 numFrames; per frame int64 timestamp, int32 depthSize, int32 imageSize, the depth bytes
(zlib stream or raw), the image bytes.  CPU only: an independent struct / zlib parse of the written file."""
import struct
import zlib

import numpy as np


def test_written_log_parses_like_rawlogreader(tmp_path):
    from kintinuous_b200 import klg
    rng = np.random.default_rng(0)
    rows, cols = 12, 16
    frames = [(1000 + 33 * k, rng.integers(0, 6000, (rows, cols)).astype(np.uint16), rng.integers(0, 256, (rows, cols, 3)).astype(np.uint8)) for k in range(3)]
    for compress in (True, False):
        p = str(tmp_path / f"x{int(compress)}.klg")
        klg.write_klg(p, frames, jpeg_encoder=lambda img: b"\xff\xd8fake-jpeg-bytes\xff\xd9", compress=compress)
        b = open(p, "rb").read()
        (n,) = struct.unpack_from("<i", b, 0)                                   # fread(&numFrames, sizeof(int32_t), 1, fp)
        assert n == 3
        off = 4
        for ts, depth, image in frames:
            t, dsz, isz = struct.unpack_from("<qii", b, off); off += 16          # timestamp, compressedDepthSize, compressedImageSize
            assert t == ts
            d = b[off:off + dsz]; off += dsz
            i = b[off:off + isz]; off += isz
            if compress:
                assert np.array_equal(np.frombuffer(zlib.decompress(d), np.uint16).reshape(rows, cols), depth)      # uncompress(), RawLogReader.cpp:108
                assert i.startswith(b"\xff\xd8")
            else:
                assert dsz == rows * cols * 2 and isz == rows * cols * 3          # the raw branches (:75, :100)
                assert np.array_equal(np.frombuffer(d, np.uint16).reshape(rows, cols), depth)
                assert np.array_equal(np.frombuffer(i, np.uint8).reshape(rows, cols, 3), image)
        assert off == len(b)
