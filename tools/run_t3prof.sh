# one CTA's per-role cycle counts of the ring kernel (-DT3_PROF build in ab_libs/lib_t3prof.so), then the shipped library's time
cp fanlin-rs_b200/libfanlin_device.so /tmp/keep.so
cp ab_libs/lib_t3prof.so fanlin-rs_b200/libfanlin_device.so
for s in ${T3_SHAPES:-c1:1024 c3:148}; do python tools/prof_resample.py ${s%%:*} ${s##*:} $T3_ARGS 2>&1 | tail -6 > gpurun_out/t3prof_${s%%:*}.log; done
cp /tmp/keep.so fanlin-rs_b200/libfanlin_device.so
for s in ${T3_SHAPES:-c1:1024 c3:148}; do python tools/prof_resample.py ${s%%:*} ${s##*:} 2>&1 | tail -1 >> gpurun_out/t3prof_${s%%:*}.log; cat gpurun_out/t3prof_${s%%:*}.log; done
