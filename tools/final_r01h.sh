#!/bin/bash
# Round-1 closing measurements (r01h): headline bench, reference arm, launch list, one full capture of the stream kernel,
# candidate-mode traces.  Run from the repository root on a B200 box (gpurun); outputs under gpurun_out/.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01h.json 2> gpurun_out/bench_r01h.err
python bench.py --impl reference --steps 4 > gpurun_out/bench_r01h_ref.json 2> gpurun_out/bench_r01h_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01h.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense > /dev/null 2> gpurun_out/ncu_l_r01h.log
ncu --set full --clock-control none --import-source on -k regex:resolve_streams -s 10 -c 1 -f -o gpurun_out/resolve_streams_r01h \
    python bench.py --no-cpu-baseline --no-dense --steps 2 > /dev/null 2> gpurun_out/ncu_rs_r01h.log
for c in c3 c4; do python tools/bench_trace.py --config $c --ref-births 200000 > gpurun_out/trace_r01h_$c.json 2> gpurun_out/trace_r01h_$c.err; done
python tools/bench_trace.py --config c5 --ref-births 200000 > gpurun_out/trace_r01h_c5.json 2> gpurun_out/trace_r01h_c5.err
echo done
