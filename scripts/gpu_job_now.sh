T="timeout 500"
$T python -m pytest tests -m gpu -q 2>&1 | tail -5
$T python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_now.err | tail -n 1 > gpurun_out/bench_now.json; cut -c1-200 gpurun_out/bench_now.json
