keep=/tmp/keep_$$.so
cp fanlin-rs_b200/libfanlin_device.so $keep
for rep in 1 2; do
  for lib in ab_libs/lib_head.so ab_libs/lib_new.so; do
    cp $lib fanlin-rs_b200/libfanlin_device.so
    for a in "c3 148" "c1crop 1024"; do echo -n "$lib "; timeout 100 python tools/prof_resample.py $a 2>&1 | tail -1; done
  done
done
cp $keep fanlin-rs_b200/libfanlin_device.so
