timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>gpurun_out/b1.err | cut -c1-180
timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b1.err | cut -c1-180
timeout 300 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/b1.err | cut -c1-180
