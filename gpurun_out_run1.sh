mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log )
tail -n 30 gpurun_out/pytest.log
