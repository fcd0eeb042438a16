set -x
cd /root/repo
timeout 600 python bench.py > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "bench rc=$?" >> gpurun_out/bench12.err
