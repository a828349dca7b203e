set -u
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_bands.py -x -q -m gpu 2>&1 | tail -3
$TR --master-port 29601 bench_bands.py --gpus $N --steps 3 --check > gpurun_out/bands_r2c_n$N.json 2> gpurun_out/bands_r2c_n$N.err; echo "bands n$N rc=$?"
$TR --master-port 29602 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c_c2_n$N.json 2> gpurun_out/bench_r2c_c2_n$N.err; echo "c2 n$N rc=$?"
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bands_r2c_n$N.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['encode'], d['decode']['ms'], d['matches_one_band_encode'])
d=json.loads([l for l in open('gpurun_out/bench_r2c_c2_n$N.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['config']['step'][:60], d['config']['single_stream_ms_per_step'], d['config']['multi_stream_ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['north_star_kernel'])"
tail -n 5 gpurun_out/bands_r2c_n$N.err gpurun_out/bench_r2c_c2_n$N.err
