set -u
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
free -g | head -2
python bench_bands.py --steps 3 > gpurun_out/bands_r2_n1.json 2> gpurun_out/bands_r2_n1.err; echo "bands n1 rc=$?"
$TR --master-port 29601 bench_bands.py --gpus $N --steps 3 --check > gpurun_out/bands_r2_n$N.json 2> gpurun_out/bands_r2_n$N.err; echo "bands n$N rc=$?"
$TR --master-port 29602 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2_c2_n$N.json 2> gpurun_out/bench_r2_c2_n$N.err; echo "c2 n$N rc=$?"
$TR --master-port 29603 bench.py --gpus $N --config c4 --steps 3 --warmup 3 > gpurun_out/bench_r2_c4_n$N.json 2> gpurun_out/bench_r2_c4_n$N.err; echo "c4 n$N rc=$?"
$TR --master-port 29604 bench.py --gpus $N --config c3 --steps 2 --warmup 3 > gpurun_out/bench_r2_c3_n$N.json 2> gpurun_out/bench_r2_c3_n$N.err; echo "c3 n$N rc=$?"
free -g | head -2
for f in gpurun_out/bands_r2_n1.json gpurun_out/bands_r2_n$N.json; do python -c "
import json,sys
d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print('$f', d['ms_per_step'], d['encode'], d['decode']['ms'], d['matches_one_band_encode'])"; done
for f in gpurun_out/bench_r2_c2_n$N.json gpurun_out/bench_r2_c4_n$N.json gpurun_out/bench_r2_c3_n$N.json; do python -c "
import json,sys
d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); e=d['e2e']; print('$f', round(d['ms_per_step'],2), round(d['value']), round(e['value']), round(e['ms_per_step'],1), e.get('copy_floor_ms'), d['parity'].get('failures_all_ranks'))"; done
tail -3 gpurun_out/*_r2_*n$N.err | tail -30
