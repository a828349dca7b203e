#!/bin/bash
# Round-end GPU job: full GPU test suite, the bench lines, the ncu launch list and one --set full capture
# of the main kernels of a C2 step (each ncu pass only after its command exited 0 without ncu).
set -u
TAG=${TAG:-r2f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err; echo "bench c2 rc=$?"
python bench.py --impl reference --steps 2 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "bench ref rc=$?"
for c in c1 c3 c4; do python bench.py --config $c --steps 3 > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err; echo "bench $c rc=$?"; done
python bench_bands.py --steps 3 --check > gpurun_out/bands_${TAG}_n1.json 2> gpurun_out/bands_${TAG}_n1.err; echo "bands rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c2_s2.json 2> gpurun_out/bench_${TAG}_c2_s2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${TAG}.log 2>&1; echo "launch list rc=$?"
HIC_ENTROPY_SERIAL=1 python tools/step_once.py 1024 426 640 2 > gpurun_out/step_once_${TAG}.log 2>&1 && \
HIC_ENTROPY_SERIAL=1 ncu --set full --clock-control none --import-source on \
    -k regex:"forward_kernel|tie_list|fixup_kernel|rle_tile_summary|rle_emit|dc_diff|pack_tile_bits|pack_emit|build_tables|huffman_sync|huffman_write|expand_tile_sum|expand_scatter|dc_prefix|inverse_kernel|upsample" \
    --launch-skip 18 --launch-count 18 -o gpurun_out/prof_c2_kernels_${TAG} -f python tools/step_once.py 1024 426 640 2 > gpurun_out/ncu_full_${TAG}.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_${TAG}.log
