mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest6.log
timeout 120 python scripts/tiled_time.py > gpurun_out/r02b_tiled_v5.log 2>&1
timeout 120 python __graft_entry__.py --smoke > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02b_smoke.log
timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?" >> gpurun_out/r02b_bench.err
