mkdir -p gpurun_out
( NOSH_B200_AMG_TIMING=1 timeout 300 python profiles/amg_setup_probe.py 200 > gpurun_out/amg_probe_a.json 2> gpurun_out/amg_probe_a.err )
( NOSH_B200_AMG_PREGROW_MB=0 NOSH_B200_AMG_TIMING=1 timeout 300 python profiles/amg_setup_probe.py 200 > gpurun_out/amg_probe_b.json 2> gpurun_out/amg_probe_b.err )
( NOSH_B200_AMG_TIMING=1 timeout 300 python profiles/amg_setup_probe.py 200 > gpurun_out/amg_probe_c.json 2> gpurun_out/amg_probe_c.err )
( NOSH_B200_AMG_PREGROW_MB=0 timeout 300 python profiles/amg_setup_probe.py 200 > gpurun_out/amg_probe_d.json 2> gpurun_out/amg_probe_d.err )
( timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log )
cat gpurun_out/amg_probe_*.json; tail -n 5 gpurun_out/pytest.log
