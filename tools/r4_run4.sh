mkdir -p gpurun_out/r4
bash tools/r4_ab.sh imls_grid=-1 imls_grid=0 imls_grid=3
timeout 800 python -m pytest tests -x -q -m gpu > gpurun_out/r4/t4.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r4/t4.log
