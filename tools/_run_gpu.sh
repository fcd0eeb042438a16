set -x
cd /root/repo
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu15.log
timeout 300 python tools/bench_fusion.py > gpurun_out/bench_fusion_r1f.log 2>&1
