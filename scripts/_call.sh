mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spline.py -x -q > gpurun_out/r02b_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest2.log
BEAST_B200_TILED_V1=1 timeout 120 python scripts/tiled_time.py > gpurun_out/r02b_tiled_v1.log 2>&1
timeout 120 python scripts/tiled_time.py > gpurun_out/r02b_tiled_v2.log 2>&1
