#!/bin/bash
# Turn the artefacts of tools/round_end_gpu.sh (gpurun_out/*${TAG}*) into the committed profiles/ files.
set -e
TAG=${TAG:-r2f}
RND=${RND:-r02}
cd "$(dirname "$0")/.."
cp gpurun_out/launches_${TAG}.csv profiles/${RND}_launches_bench_c2.csv
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv gpurun_out/bench_${TAG}_c2_s2.json > /tmp/launch.md
python - <<PY
s = open('profiles/${RND}_launches_bench_c2.md').read()
note = s[:s.index('| kernel | launches |')]
open('profiles/${RND}_launches_bench_c2.md', 'w').write(note + open('/tmp/launch.md').read())
PY
cp gpurun_out/bench_${TAG}_c2.json profiles/${RND}_bench_c2_n1.json
cp gpurun_out/bench_${TAG}_ref.json profiles/${RND}_bench_c2_reference_arm.json
python - <<PY
import csv, subprocess, json, re
rep = 'gpurun_out/prof_c2_kernels_${TAG}.ncu-rep'
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; units = rows[1]
col = hdr.index
traffic = {}
for r in rows[2:]:
    name = r[col('Kernel Name')]
    m = re.search(r'(\w+)(<[^(]*>)?\(', name); tag = m.group(1) if m else name
    if tag == 'huffman_sync_kernel' and ('<1>' in name or '<(bool)1>' in name): tag = 'huffman_resync_kernel'
    if tag in traffic: continue
    def val(k):
        return float(r[col(k)].replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[col(k)], 1)
    t = float(r[col('gpu__time_duration.sum')]); tu = units[col('gpu__time_duration.sum')]
    traffic[tag] = {'dram_read_bytes': val('dram__bytes_read.sum'), 'dram_write_bytes': val('dram__bytes_write.sum'),
                    'ms_under_ncu': t if tu == 'ms' else t / 1e3,
                    'issue_active_pct': float(r[col('smsp__issue_active.avg.pct_of_peak_sustained_active')]),
                    'dram_pct_of_peak': float(r[col('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')]),
                    'registers': int(float(r[col('launch__registers_per_thread')])), 'kernel': name[:90]}
    print('%-28s %8.3f ms  rd %7.1f MB wr %7.1f MB  issue %5.1f%%' % (tag, traffic[tag]['ms_under_ncu'],
          traffic[tag]['dram_read_bytes'] / 1e6, traffic[tag]['dram_write_bytes'] / 1e6, traffic[tag]['issue_active_pct']))
json.dump(traffic, open('profiles/${RND}_traffic_c2.json', 'w'), indent=1)
PY
python tools/ncu_summary.py gpurun_out/prof_c2_kernels_${TAG}.ncu-rep "Round 2 (final) -- ncu --set full --clock-control none of the main kernels of one C2 step (tools/step_once.py 1024 426 640 2 with HIC_ENTROPY_SERIAL=1, second step captured)" > profiles/${RND}_c2_kernels_ncu.md
