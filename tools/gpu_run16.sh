timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad" 2>&1 | tail -5
timeout 300 python tools/bench_kernels.py --iters 10 2>&1 | grep wgrad | head -3 | cut -c1-200
