set -x
cd /root/repo
timeout 600 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?" >> gpurun_out/bench13.err
for T in 1 10; do
for DT in bf16 f32; do
MILB200_TAPE_GRAPHS=0 timeout 300 python tools/profile_fusion.py 15592 $T $DT 2 > gpurun_out/plain_pf_${T}_${DT}.log 2>&1 && \
MILB200_TAPE_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_fusion_r1f_T${T}_${DT}.csv python tools/profile_fusion.py 15592 $T $DT 2 > gpurun_out/ncu_pf_${T}_${DT}.log 2>&1
done; done
