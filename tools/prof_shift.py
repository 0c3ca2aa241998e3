#!/usr/bin/env python
"""Workload for ncu captures of the kernels that only run on a shift frame (extract_kernel, clear_planes_*) and of the GUI taps:
40 frames with a 2-voxel shift threshold (a +x shift every ~5 frames), then finalise (full-volume extraction) and one live image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kintinuous_b200 as kb
from kintinuous_b200 import synth
vol = int(sys.argv[1]) if len(sys.argv) > 1 else 512
t = kb.Tracker(kb.Config.default(vol=vol, voxel_shift=2))
t.set_slice_processing(True, 8)          # CloudSliceProcessor on the device (kt_slice.cu) for every slice, the FINAL one included
fr = [synth.render(k) for k in range(24)]
for k in range(24):
    t.process_frame(fr[k][0], fr[k][1], k)
t.live_image()
t.finalise()
print("slices", t.num_slices(), "launches", t.launch_count())
