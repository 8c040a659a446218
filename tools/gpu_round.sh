#!/bin/bash
# One full measurement pass on the GPU box: parity tests, bench (ours + reference arm), ncu launch lists, ncu --set full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
# The ncu passes run the host-driven LM loop (PPO_BA_NO_GRAPH=1): same kernels as the captured graph, one launch per node.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"
export PPO_BA_NO_GRAPH=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv python tools/profile_one.py 2 > $OUT/${TAG}_prof_launch.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches_config4.csv python tools/profile_one.py 4 > $OUT/${TAG}_prof_launch4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_point_linearize|k_pose_accumulate|k_plane_jac|k_schur_bd|k_schur_pairs|k_backsub|k_backsub_points|k_point_residual|k_cuboid_jac|k_update" -c 12 -o $OUT/${TAG}_ncu_assembly python tools/profile_one.py 2 > $OUT/${TAG}_prof_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_chol_dataflow|k_backsolve_chain" -c 4 -o $OUT/${TAG}_ncu_solve python tools/profile_one.py 2 > $OUT/${TAG}_prof_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_chol_dataflow|k_backsolve_chain|k_schur_pairs|k_point_linearize" -c 4 -o $OUT/${TAG}_ncu_config4 python tools/profile_one.py 4 > $OUT/${TAG}_prof_c.log 2>&1
unset PPO_BA_NO_GRAPH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/${TAG}_smi.csv
ls -la $OUT | tail -15
