#!/bin/bash
# One GPU session collecting the numbers and profiles committed under profiles/ (round 2).
# usage (under gpurun): bash scripts/collect_evidence.sh
set -x
O=gpurun_out/r2
mkdir -p $O
python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for c in 0 2 3 4; do python bench.py --config $c --steps 30 --warmup 5 > $O/bench_config$c.json 2> $O/bench_config$c.err; done
python bench.py --config 4 --rois-per-img 8192 --steps 10 --warmup 3 > $O/bench_config4_K8192.json 2> $O/bench_config4_K8192.err
python bench.py --nchw --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs > $O/bench_nchw.json 2> $O/bench_nchw.err
for k in 512 1024 2048 4096 8192; do PROBE_K=$k PROBE_B=1 python scripts/roi_probe.py; PROBE_BF16=1 PROBE_K=$k PROBE_B=1 python scripts/roi_probe.py; done > $O/roi_k_sweep.txt 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --no-verify"
$B > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv $B > $O/ncu_launches.log 2>&1
python scripts/roi_probe.py ncu > $O/plain_roi.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"roi_fuse_fwd_ring|roi_bwd_pull_tma|roi_bin_kernel|roi_prep" -s 8 -c 4 -o $O/roi_final python scripts/roi_probe.py ncu > $O/ncu_roi.log 2>&1
python scripts/fpn_bwd_probe.py 3 > $O/plain_fpn.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"fpn_bwd_fused_tma|gather_bwd_up" -s 2 -c 2 -o $O/fpn_bwd_final python scripts/fpn_bwd_probe.py 3 > $O/ncu_fpn.log 2>&1
tail -2 $O/ncu_roi.log $O/ncu_fpn.log $O/ncu_launches.log
