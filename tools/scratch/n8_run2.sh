set -u
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 bench_bands.py --gpus $N --steps 3 --check > gpurun_out/bands_r2d_n8.json 2> gpurun_out/bands_r2d_n8.err; echo "bands n8 rc=$?"
timeout 400 $TR --master-port 29602 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2d_c2_n8.json 2> gpurun_out/bench_r2d_c2_n8.err; echo "c2 n8 rc=$?"
timeout 500 $TR --master-port 29603 bench.py --gpus $N --config c3 --steps 2 --warmup 3 > gpurun_out/bench_r2d_c3_n8.json 2> gpurun_out/bench_r2d_c3_n8.err; echo "c3 n8 rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bands_r2d_n8.json') if l.startswith('{')][-1]); print('bands', d['ms_per_step'], d['encode'], d['decode']['ms'], d['matches_one_band_encode'])
for f in ['gpurun_out/bench_r2d_c2_n8.json','gpurun_out/bench_r2d_c3_n8.json']:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); e=d['e2e']; print(f, round(d['ms_per_step'],2), round(d['value']), d['config'].get('single_stream_ms_per_step'), d['config'].get('multi_stream_ms_per_step'), round(e['value']), round(e['ms_per_step'],1), e.get('copy_floor_ms'), d['parity'].get('failures_all_ranks'))
"
tail -n 3 gpurun_out/bench_r2d_c3_n8.err
