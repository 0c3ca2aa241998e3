"""ctypes harness of the .klg reader (kt_klg_*, include/kintinuous_b200.h) -- test / bench plumbing, like binding.py."""
from __future__ import annotations

import ctypes as C
import struct
import zlib

import numpy as np

from .binding import load, _check, Pose


class KlgFrame(C.Structure):
    _fields_ = [("timestamp", C.c_int64), ("depth_size", C.c_int32), ("image_size", C.c_int32), ("is_compressed", C.c_int), ("frame", C.c_int),
                ("depth_dev", C.c_void_p), ("rgb_dev", C.c_void_p), ("depth_host", C.c_void_p),
                ("compressed_depth", C.c_void_p), ("compressed_image", C.c_void_p)]


class KlgReader:
    def __init__(self, path, rows=480, cols=640, device=0):
        self.lib = load()
        self.rows, self.cols = rows, cols
        self.h = C.c_void_p()
        _check(self.lib.kt_klg_open(path.encode(), rows, cols, device, C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.kt_klg_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_frames(self):
        return int(self.lib.kt_klg_num_frames(self.h))

    def has_more(self):
        return bool(self.lib.kt_klg_has_more(self.h))

    def set_flip_colors(self, flip):
        _check(self.lib.kt_klg_set_flip_colors(self.h, int(flip)))

    def read_next(self):
        f = KlgFrame()
        _check(self.lib.kt_klg_read_next(self.h, C.byref(f)))
        _check(self.lib.kt_klg_wait(self.h))
        return f

    def track_next(self, tracker):
        p = Pose()
        _check(self.lib.kt_klg_track_next(self.h, tracker.h, C.byref(p)))
        return p


def write_klg(path, frames, jpeg_encoder=None, compress=True):
    """Write a .klg file the way the reference's logger does (layout from RawLogReader.cpp:29, :54-66).
    frames: iterable of (timestamp, depth u16 [rows, cols], image u8 [rows, cols, 3] in the byte order to be stored).
    compress: zlib depth + JPEG image (jpeg_encoder(image) -> bytes), else raw bytes."""
    frames = list(frames)
    with open(path, "wb") as fp:
        fp.write(struct.pack("<i", len(frames)))
        for ts, depth, image in frames:
            d = np.ascontiguousarray(depth, dtype=np.uint16).tobytes()
            i = np.ascontiguousarray(image, dtype=np.uint8).tobytes()
            if compress:
                d = zlib.compress(d)
                i = jpeg_encoder(image)
            fp.write(struct.pack("<qii", int(ts), len(d), len(i)))
            fp.write(d)
            fp.write(i)
