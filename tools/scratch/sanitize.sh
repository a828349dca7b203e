set -u
python -m pytest tests/test_gpu_wavelet.py tests/test_gpu_full_size.py -x -q -m gpu -k "wavelet or c4" 2>&1 | tail -2
python bench.py --config c4 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c4', d['ms_per_step'], d['kernels']['stream_scan64_kernel'], d['parity'])"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/step_once.py 8 426 640 1 > gpurun_out/sanitize_step.log 2>&1; echo "memcheck step rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/sanitize_step.log | head -8
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_dct.py tests/test_gpu_restarts.py tests/test_gpu_wavelet_settings.py tests/test_gpu_bands.py -x -q -m gpu -k "not 1080 and not 720 and not 2048" > gpurun_out/sanitize_tests.log 2>&1; echo "memcheck tests rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|passed|failed" gpurun_out/sanitize_tests.log | head -8
