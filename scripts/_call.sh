mkdir -p gpurun_out
BEAST_B200_TILED_ENC_CTAS=2 timeout 300 python -m pytest tests/test_gpu_spline.py -x -q > gpurun_out/r02b_pytest8b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest8b.log
timeout 120 python scripts/tiled_time.py > gpurun_out/r02b_tiled_v7a.log 2>&1
BEAST_B200_TILED_ENC_CTAS=2 timeout 120 python scripts/tiled_time.py > gpurun_out/r02b_tiled_v7b.log 2>&1
