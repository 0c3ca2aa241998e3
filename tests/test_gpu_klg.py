"""The .klg reader (kt_klg.cu: RawLogReader.cpp:20-133 + TrackerInterface.cpp:82-104) on the GPU box.

Checker: Python's zlib / struct for the container, cv2.imdecode for the JPEG (the reference calls OpenCV's cvDecodeImage, i.e. libjpeg;
this image has cv2 4.13 with libjpeg-turbo).  Bars: timestamps, sizes, flags and the DEPTH are exact; a raw image is exact; a decoded
JPEG comes from two implementations of the same standard (nvJPEG on the device, libjpeg-turbo on the host), which differ in IDCT rounding
and -- for the usual 4:2:0 files -- in how the half-resolution chroma planes are interpolated back (libjpeg's "fancy" triangle filter
vs replication).  Measured on the synthetic frames, whose 7.85 cm colour checker is the worst case for chroma interpolation: 4:2:0 mean
|d| 0.87 grey levels, 99.9 % within 8, single bytes up to 22 on checker edges; the bars below are mean < 1.5 and 99.9 % <= 12 for 4:2:0,
mean < 0.7 and 99.9 % <= 4 for a 4:4:4 file (no chroma interpolation: what remains is IDCT rounding)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _d2h(ptr, shape, dtype):
    """device pointer -> numpy (cudaMemcpy through libcudart, the runtime the product links)"""
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    out = np.empty(shape, dtype)
    assert rt.cudaMemcpy(C.c_void_p(out.ctypes.data), C.c_void_p(ptr), C.c_size_t(out.nbytes), 2) == 0
    return out


def _frames(n, rows, cols):
    from kintinuous_b200 import synth
    return [(1_000_000 + 33_333 * k,) + synth.render(k, cols, rows) for k in range(n)]


def test_klg_reader_compressed_and_raw(built, tmp_path):
    import cv2
    from kintinuous_b200 import klg
    rows, cols = 240, 320
    fr = _frames(5, rows, cols)
    enc420 = lambda img: cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes()
    enc444 = lambda img: cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])[1].tobytes()
    pc, p4, pr = str(tmp_path / "c.klg"), str(tmp_path / "c444.klg"), str(tmp_path / "r.klg")
    klg.write_klg(pc, fr, jpeg_encoder=enc420, compress=True)
    klg.write_klg(p4, fr, jpeg_encoder=enc444, compress=True)
    klg.write_klg(pr, fr, compress=False)
    for path, compressed, enc, bar in ((pc, True, enc420, (1.5, 12)), (p4, True, enc444, (0.7, 4)), (pr, False, None, None)):
        rd = klg.KlgReader(path, rows, cols)
        assert rd.num_frames == 5
        for k in range(5):
            f = rd.read_next()
            assert f.timestamp == fr[k][0] and f.frame == k + 1 and bool(f.is_compressed) == compressed
            import ctypes as C
            assert np.array_equal(_d2h(f.depth_dev, (rows, cols), np.uint16), fr[k][1]), k      # depth: exact
            host_depth = np.ctypeslib.as_array(C.cast(f.depth_host, C.POINTER(C.c_uint16)), shape=(rows, cols))
            assert np.array_equal(host_depth, fr[k][1])
            got = _d2h(f.rgb_dev, (rows, cols, 3), np.uint8)
            if not compressed:
                assert np.array_equal(got, fr[k][2]), k
                assert f.depth_size == rows * cols * 2 and f.image_size == rows * cols * 3
            else:
                want = cv2.imdecode(np.frombuffer(enc(fr[k][2]), np.uint8), cv2.IMREAD_COLOR)    # interleaved B,G,R like cvDecodeImage
                diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
                print(f"frame {k}: JPEG nvJPEG vs libjpeg-turbo: mean |d| {diff.mean():.3f}, 99.9 % {np.quantile(diff, 0.999):.0f}, max {diff.max()}")
                assert diff.mean() < bar[0] and np.quantile(diff, 0.999) <= bar[1], bar
            assert rd.has_more() == (k + 1 < 4)                                                  # currentFrame + 1 < numFrames
        rd.close()
    # -f: channels swapped on the device
    rd = klg.KlgReader(pr, rows, cols); rd.set_flip_colors(True)
    f = rd.read_next()
    assert np.array_equal(_d2h(f.rgb_dev, (rows, cols, 3), np.uint8), fr[0][2][..., ::-1])
    rd.close()


def test_klg_track_next_equals_feeding_the_frames(built, tmp_path):
    """TrackerInterface::process through the reader == the same frames handed to kt_process_frame: identical poses (raw log: bit for bit)."""
    import kintinuous_b200 as kb
    from kintinuous_b200 import klg
    rows, cols = 240, 320
    fr = _frames(6, rows, cols)
    p = str(tmp_path / "t.klg")
    klg.write_klg(p, fr, compress=False)
    cfg = kb.Config.default(rows=rows, cols=cols, vol=128, odometry=0)
    a = kb.Tracker(cfg); b = kb.Tracker(cfg)
    rd = klg.KlgReader(p, rows, cols)
    for k in range(6):
        pa = rd.track_next(a)
        pb = b.process_frame(fr[k][1], fr[k][2], fr[k][0])
        for x, y in zip(pa.as_tuple(), pb.as_tuple()):
            assert np.array_equal(x, y), k
    rd.close(); a.close(); b.close()
