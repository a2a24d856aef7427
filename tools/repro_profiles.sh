#!/bin/bash
# Reproduces the evidence under profiles/ on a B200 box (run from the repository root, e.g. through gpurun).
# Every number quoted in DESIGN.md comes from one of these commands; ncu captures are taken only after the
# same command has run clean without the profiler, and no benchmark value is ever taken from a run under ncu.
set -e
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q
python bench.py > gpurun_out/bench.json                         # headline line (C2, e2e + resident + roofline + cpu_baseline + dense side measurement)
python bench.py --impl reference > gpurun_out/bench_ref.json    # the unmodified reference on the same simulations
for c in c2x c3 c4; do python tools/bench_dense.py --config $c > gpurun_out/dense_${c}_full.json; done
for c in c3 c4 c5; do python tools/bench_trace.py --config $c > gpurun_out/trace_$c.json; done
# launch list of the bench (kernel shares) and full captures of the two top kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense > /dev/null
# (-s 66: the last launches of the run are the steady steps, grid 1000; earlier ones belong to the set-up generations)
ncu --set full --clock-control none --import-source on -k regex:resolve_streams -s 66 -c 1 -f -o gpurun_out/resolve_streams \
    python bench.py --no-cpu-baseline --no-dense --steps 2 > /dev/null
for c in c2x c3 c4; do
  ncu --set full --clock-control none --import-source on -k regex:classify_dense_tiled -s 1 -c 1 -f -o gpurun_out/dense_$c \
      python tools/bench_dense.py --config $c --births 65536 --check 0 > /dev/null
done
# text summaries as committed under profiles/: python tools/ncu_summary.py launches gpurun_out/launches.csv ; python tools/ncu_summary.py kernel gpurun_out/resolve_streams.ncu-rep
# per-stream timeline of the stream kernel
APT_STREAM_PROF=gpurun_out/stream_prof.txt python bench.py --no-cpu-baseline --no-dense --steps 2 > /dev/null
# multi-GPU (needs 2 / 8 GPUs): fused peer-memory exchange vs NCCL all-gather, replica scaling of the tree bench
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_dense.py --config c4 --exchange p2p
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 bench.py --gpus 8
