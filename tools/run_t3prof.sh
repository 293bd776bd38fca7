cp fanlin-rs_b200/libfanlin_device.so /tmp/keep.so
cp ab_libs/lib_t3prof.so fanlin-rs_b200/libfanlin_device.so
python tools/prof_resample.py c1 1024 2>&1 | tail -6 > gpurun_out/t3prof_c1.log
python tools/prof_resample.py c3 148 2>&1 | tail -6 > gpurun_out/t3prof_c3.log
cp /tmp/keep.so fanlin-rs_b200/libfanlin_device.so
python tools/prof_resample.py c1 1024 2>&1 | tail -1 >> gpurun_out/t3prof_c1.log
python tools/prof_resample.py c3 148 2>&1 | tail -1 >> gpurun_out/t3prof_c3.log
cat gpurun_out/t3prof_c1.log gpurun_out/t3prof_c3.log
