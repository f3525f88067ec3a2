mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spline.py tests/test_gpu_bpe.py -x -q > gpurun_out/r02b_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest5.log
timeout 120 python scripts/tiled_time.py > gpurun_out/r02b_tiled_v4.log 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tiled -s 7 -c 2 -f -o gpurun_out/r02b_tiled python scripts/profile_shipped.py > gpurun_out/r02b_tiled_ncu.log 2>&1
