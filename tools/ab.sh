#!/bin/bash
# A/B of built libraries on one GPU box: tools/ab.sh libA.so libB.so ... (each run AB_REPS times, interleaved).
# AB_STEPS (default 25) x 4096 C2 images reaches the power cap the driver's 20-step run sees.
keep=/tmp/keep_$$.so
cp fanlin-rs_b200/libfanlin_device.so $keep
for rep in ${AB_REPS:-1 2}; do
  for lib in "$@"; do
    cp "$lib" fanlin-rs_b200/libfanlin_device.so
    python bench.py --steps ${AB_STEPS:-25} --warmup 5 --no-e2e --no-cpu-baseline --no-parity --no-configs ${AB_ARGS} 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['roofline']['kernel'], round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks'].get('sm_mhz_min_under_load'), d['clocks'].get('power_w_max'), d['clocks']['reasons'])"
  done
done
cp $keep fanlin-rs_b200/libfanlin_device.so
