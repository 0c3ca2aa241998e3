for c in 12 16 24 32; do KT_INT_ZCHUNKS=$c python tools/stage_ab.py 30 512 0 2>&1 | tail -1; done
for c in 16 24 32 48; do KT_INT_ZCHUNKS=$c python tools/stage_ab.py 24 1024 0 2>&1 | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slice_normals_kernel -s 0 -c 1 -f -o gpurun_out/prof_r2v2_slice_normals_kernel python tools/prof_shift.py > gpurun_out/ncu_r2v2_slice_normals_kernel.log 2>&1; tail -1 gpurun_out/ncu_r2v2_slice_normals_kernel.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slice_accumulate_kernel -s 0 -c 1 -f -o gpurun_out/prof_r2v2_slice_accumulate_kernel python tools/prof_shift.py > gpurun_out/ncu_r2v2_slice_accumulate_kernel.log 2>&1; tail -1 gpurun_out/ncu_r2v2_slice_accumulate_kernel.log
python tools/icp_prof.py 2>&1 | tail -25
