mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_mirror.py -m gpu -x -q -k "arclength or cpp or mirror or continuation" > gpurun_out/arc_check.log 2>&1
tail -15 gpurun_out/arc_check.log
