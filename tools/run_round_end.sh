# Round-end records on one GPU box: the driver's bench command, then (only after it exited 0) the ncu launch list and one
# --set full capture of the ring kernel on C3.  Numbers printed under ncu are never bench values.
set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err || exit 1
tail -c 600 gpurun_out/bench_r02b.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_r02b.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --batch 1184 --only-configs C1,C3,C5_crop > gpurun_out/ncu_launches_r02b.log 2>&1
python tools/prof_resample.py c3 148 > gpurun_out/prof_c3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_resample_tc3 -c 1 -o gpurun_out/ncu_tc3_r02b -f python tools/prof_resample.py c3 148 > gpurun_out/ncu_tc3_r02b.log 2>&1
python tools/prof_resample.py c1 1024 > gpurun_out/prof_c1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_resample_tc3 -c 1 -o gpurun_out/ncu_tc3_c1_r02b -f python tools/prof_resample.py c1 1024 > gpurun_out/ncu_tc3_c1_r02b.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
python tools/prof_blur.py 148 > gpurun_out/prof_blur_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:blur_tc2 -c 1 -o gpurun_out/ncu_blur_r02b -f python tools/prof_blur.py 148 > gpurun_out/ncu_blur_r02b.log 2>&1
python tools/latency.py > gpurun_out/latency.jsonl 2> gpurun_out/latency.err
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
