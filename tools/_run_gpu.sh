set -x
cd /root/repo
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu17.log
