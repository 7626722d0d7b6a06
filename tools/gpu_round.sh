#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list (+ DRAM traffic), ncu --set full of the top kernels. Outputs under gpurun_out/<tag>_*.
tag=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/${tag}_bench.log | cut -c1-1500
timeout 300 python bench.py --workload smmnist --no-cpu-baseline > gpurun_out/${tag}_bench_smmnist.log 2>&1; echo "bench smmnist rc=$?"; tail -1 gpurun_out/${tag}_bench_smmnist.log | cut -c1-600
timeout 300 python bench.py --impl reference --steps 2 > gpurun_out/${tag}_bench_reference.log 2>&1; echo "bench reference rc=$?"; tail -1 gpurun_out/${tag}_bench_reference.log | cut -c1-600
timeout 300 python bench.py --workload human_rollout > gpurun_out/${tag}_bench_rollout.log 2>&1; echo "bench rollout rc=$?"; tail -1 gpurun_out/${tag}_bench_rollout.log | cut -c1-600
timeout 200 python tools/step_timeline.py > gpurun_out/${tag}_timeline.log 2>&1; echo "timeline rc=$?"; head -1 gpurun_out/${tag}_timeline.log
timeout 200 python tools/step_profile.py --top 70 > gpurun_out/${tag}_step_profile.log 2>&1; echo "step profile rc=$?"; head -1 gpurun_out/${tag}_step_profile.log
timeout 200 python tests/dev_thin.py > gpurun_out/${tag}_thin.log 2>&1; echo "thin rc=$?"
if [ "$2" != "noncu" ]; then
# every launch of ~1.5 steps with its device time and DRAM traffic (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1300 -c 700 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
for c in d30 e22 d00; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_conv3x3_$c \
    python tests/dev_prof_conv.py $c > gpurun_out/${tag}_ncu_conv_$c.log 2>&1; echo "ncu conv $c rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad3x3 -s 2 -c 1 -f -o gpurun_out/${tag}_wgrad3x3 \
    python tests/dev_prof_wgrad.py > gpurun_out/${tag}_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_bwd -s 4 -c 4 -f -o gpurun_out/${tag}_bn_bwd \
    python tests/dev_bn_prof.py > gpurun_out/${tag}_ncu_bnbwd.log 2>&1; echo "ncu bn_bwd rc=$?"
# the HBM-bound ends: thin-input convolution, thin weight gradient (both operand roles), decoder head
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:thin_|decoder_head' -s 8 -c 6 -f -o gpurun_out/${tag}_thin \
    python tests/dev_thin.py > gpurun_out/${tag}_ncu_thin.log 2>&1; echo "ncu thin rc=$?"
fi
