#!/bin/bash
# One gpurun call: GPU tests, the default bench line, the x-order switch A/B, host-core emulation, the eps x minPts
# sweep (BASELINE configs[4]) and the ncu captures of this round.  Everything lands in gpurun_out/a/.
O=gpurun_out/a
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }

stamp tests
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
stamp bench_default
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "rc=$?" >> $O/bench_c2.err
stamp bench_xgrid0
CL2_XGRID=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-exact --no-files > $O/bench_c2_xgrid0.json 2> $O/bench_c2_xgrid0.err
stamp hostcores
for cfg in "8 2" "4 0" "4 2" "8 1"; do
  set -- $cfg
  CL2_THREADS=$1 CL2_BLOCKING_SYNC=$2 timeout 200 python bench.py --host-cores 4 --steps 5 --warmup 3 --no-cpu-baseline --no-exact \
      --no-files --no-kernel-timing > $O/bench_hc4_t$1_m$2.json 2> $O/bench_hc4_t$1_m$2.err
done
stamp sweep
timeout 300 python profiles/prof_sweep.py 0 > $O/r2_sweep.md 2> $O/r2_sweep.err; echo "rc=$?" >> $O/r2_sweep.err
stamp ncu_launches
export CL2_NO_POOL=1
python profiles/prof_target.py 8.06e6 2 > $O/prof_target_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_prof_target.csv \
    python profiles/prof_target.py 8.06e6 2 > $O/prof_target_ncu.log 2>&1
stamp ncu_full_target
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_xorder_from_grid|k_est_bg$|k_block_adj$' -c 6 \
    -f -o $O/prof_target_full python profiles/prof_target.py 8.06e6 1 > $O/prof_target_full.log 2>&1
stamp ncu_full_hic
python profiles/prof_hic_target.py 3e6 > $O/prof_hic_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_c2_core|k_c2_edges|k_c2_union|k_block_union|k_block_adj_heavy|k_c2_border_mask' -c 10 \
    -f -o $O/prof_hic_full python profiles/prof_hic_target.py 3e6 > $O/prof_hic_full.log 2>&1
stamp done
