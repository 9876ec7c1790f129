mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_full.log 2>&1; echo rc=$? >> gpurun_out/pytest_full.log )
grep -E "passed|failed|GPU .* iterations|Newton n=|true relative|SHIM" gpurun_out/pytest_full.log | tail -14
