set -x
for dbg in 0 1 2; do HIC_K1_DEBUG=$dbg python tools/quick_time.py 2>&1 | grep forward; done
for dbg in 0 1 2; do HIC_LIB_PATH=/root/repo/tools/build/lib_th32.so HIC_K1_DEBUG=$dbg python tools/quick_time.py 2>&1 | grep forward; done
HIC_K1_BPT=2 python tools/quick_time.py 2>&1 | grep forward
python tools/step_once.py 1024 426 640 3 2>&1 | tail -40
