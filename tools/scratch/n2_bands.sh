set -u
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_bands.py -x -q -m gpu 2>&1 | tail -3
python bench_bands.py --steps 3 > gpurun_out/bands_r2b_n1.json 2> gpurun_out/bands_r2b_n1.err; echo "bands n1 rc=$?"
$TR --master-port 29601 bench_bands.py --gpus $N --steps 3 --check > gpurun_out/bands_r2b_n$N.json 2> gpurun_out/bands_r2b_n$N.err; echo "bands n$N rc=$?"
for f in gpurun_out/bands_r2b_n1.json gpurun_out/bands_r2b_n$N.json; do python -c "
import json,sys
d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print('$f', d['ms_per_step'], d['encode'], d['decode']['ms'], d['matches_one_band_encode'])"; done
tail -n 5 gpurun_out/bands_r2b_n$N.err
