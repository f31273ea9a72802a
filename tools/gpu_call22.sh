mkdir -p gpurun_out/r2
timeout 180 python -m pytest tests/test_sdf_build.py -x -q -m gpu -k "tensor_core and c1" -s > gpurun_out/r2/t22_umma.log 2>&1; echo umma c1 rc=$?; tail -12 gpurun_out/r2/t22_umma.log
