set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_feeder.py -x -q -m gpu > gpurun_out/pytest_gpu14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu14.log
for L in 0 1; do
MILB200_TAPE_LANES=$L timeout 300 python tools/bench_fusion.py > gpurun_out/bench_fusion_lanes$L.log 2>&1
done
