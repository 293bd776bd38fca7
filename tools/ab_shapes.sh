# A/B of two built libraries (ab_libs/lib_head.so, ab_libs/lib_new.so) on resample shapes: AB_SHAPES="c5:592 c2crop:1184"
keep=/tmp/keep_$$.so
cp fanlin-rs_b200/libfanlin_device.so $keep
for rep in 1 2; do
  for lib in ab_libs/lib_head.so ab_libs/lib_new.so; do
    cp $lib fanlin-rs_b200/libfanlin_device.so
    for s in ${AB_SHAPES:-c5:592 c2crop:1184}; do echo -n "$lib "; timeout 100 python tools/prof_resample.py ${s%%:*} ${s##*:} 2>&1 | tail -1; done
  done
done
cp $keep fanlin-rs_b200/libfanlin_device.so
