for v in 0 1 3 4; do echo "pf waves $v"; HIC_LIB_PATH=/root/repo/tools/build/lib_pf$v.so python tools/step_once.py 1024 426 640 3 2>&1 | grep -E "forward_kernel"; done
echo "pf waves 2 (product)"; python tools/step_once.py 1024 426 640 3 2>&1 | grep -E "forward_kernel"
python tools/step_once.py 256 2160 3840 2 2>&1 | grep -E "forward_kernel"
HIC_LIB_PATH=/root/repo/tools/build/lib_pf4.so python tools/step_once.py 256 2160 3840 2 2>&1 | grep -E "forward_kernel"
HIC_LIB_PATH=/root/repo/tools/build/lib_pf0.so python tools/step_once.py 256 2160 3840 2 2>&1 | grep -E "forward_kernel"
