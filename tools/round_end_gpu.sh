#!/bin/bash
# Round-end GPU job: full GPU test suite, the bench lines, the ncu launch list and one --set full capture
# of the main kernels of a C2 step (each ncu pass only after its command exited 0 without ncu).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1f.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu_r1f.log
python bench.py > gpurun_out/bench_r1f_c2.json 2> gpurun_out/bench_r1f_c2.err; echo "bench c2 rc=$?"
python bench.py --impl reference --steps 2 > gpurun_out/bench_r1f_ref.json 2> gpurun_out/bench_r1f_ref.err; echo "bench ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1f_c2_s2.json 2> gpurun_out/bench_r1f_c2_s2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1f.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_r1f.log 2>&1; echo "launch list rc=$?"
HIC_ENTROPY_SERIAL=1 python tools/step_once.py 1024 426 640 2 > gpurun_out/step_once_r1f.log 2>&1 && \
HIC_ENTROPY_SERIAL=1 ncu --set full --clock-control none --import-source on \
    -k regex:"forward_kernel|fixup_kernel|rle_tile_summary|rle_emit|dc_diff|pack_tile_bits|pack_emit|build_tables|huffman_sync|huffman_write|expand_tile_sum|expand_scatter|dc_write|inverse_kernel|upsample" \
    --launch-skip 17 --launch-count 17 -o gpurun_out/prof_c2_kernels_r1f -f python tools/step_once.py 1024 426 640 2 > gpurun_out/ncu_full_r1f.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_r1f.log
