set -u
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 bench_bands.py --gpus $N --steps 3 --check > gpurun_out/bands_r2e_n4.json 2> gpurun_out/bands_r2e_n4.err; echo "bands n4 rc=$?"
timeout 300 $TR --master-port 29602 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2e_c2_n4.json 2> gpurun_out/bench_r2e_c2_n4.err; echo "c2 n4 rc=$?"
timeout 300 $TR --master-port 29603 bench.py --gpus $N --config c4 --steps 3 --warmup 3 > gpurun_out/bench_r2e_c4_n4.json 2> gpurun_out/bench_r2e_c4_n4.err; echo "c4 n4 rc=$?"
timeout 400 $TR --master-port 29604 bench.py --gpus $N --config c3 --steps 2 --warmup 3 > gpurun_out/bench_r2e_c3_n4.json 2> gpurun_out/bench_r2e_c3_n4.err; echo "c3 n4 rc=$?"
timeout 200 $TR --master-port 29605 bench.py --gpus $N --config c1 --steps 5 --warmup 3 > gpurun_out/bench_r2e_c1_n4.json 2> gpurun_out/bench_r2e_c1_n4.err; echo "c1 n4 rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bands_r2e_n4.json') if l.startswith('{')][-1]); print('bands', d['ms_per_step'], d['encode']['ms'], d['decode']['ms'], d['matches_one_band_encode'])
for c in ['c2','c4','c3','c1']:
    d=json.loads([l for l in open('gpurun_out/bench_r2e_%s_n4.json'%c) if l.startswith('{')][-1]); e=d['e2e']; print(c, round(d['ms_per_step'],2), round(d['value']), round(e['value']), round(e['ms_per_step'],1), e.get('copy_floor_ms'), d['parity'].get('failures_all_ranks', d['parity'].get('mismatches')))
"
