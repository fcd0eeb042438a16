cd /root/repo
timeout 300 python tools/pyprof_fusion.py > gpurun_out/pyprof_r1f.log 2>&1
