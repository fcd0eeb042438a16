set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu18.log
timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/bench16.json 2> gpurun_out/bench16.err; echo "rc=$?" >> gpurun_out/bench16.err
