timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 800 python tools/bench_configs.py > gpurun_out/r1o_configs.json 2> gpurun_out/r1o_configs.err; grep -c "parity_ok\": true" gpurun_out/r1o_configs.json
timeout 200 python tools/bench_small_vs.py > gpurun_out/r1o_small.json 2>/dev/null; cut -c1-40,95-200 gpurun_out/r1o_small.json
TAG=r1o bash tools/gpu_sanitize_rounds.sh
