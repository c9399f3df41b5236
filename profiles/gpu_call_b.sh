#!/bin/bash
# Second gpurun call of the round: GPU tests and the default bench line with the scratch arena and the 4-rows-per-round
# x-order walk, arena A/B (step time and host CPU with sleeping waits on 4 cores), the other BASELINE configs, and the
# ncu capture of the new kernel.  Everything lands in gpurun_out/b/.
O=gpurun_out/b
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
LEAN="--no-cpu-baseline --no-exact --no-files"

stamp tests
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
stamp smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
stamp bench_default
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "rc=$?" >> $O/bench_c2.err
stamp bench_arena0
CL2_ARENA=0 timeout 300 python bench.py --steps 5 --warmup 3 $LEAN > $O/bench_c2_arena0.json 2> $O/bench_c2_arena0.err
stamp hostcores
CL2_BLOCKING_SYNC=1 timeout 200 python bench.py --host-cores 4 --steps 5 --warmup 3 $LEAN --no-kernel-timing --extra-sync-modes 2,0 \
    > $O/bench_hc4_m1_arena1.json 2> $O/bench_hc4_m1_arena1.err
CL2_ARENA=0 CL2_BLOCKING_SYNC=1 timeout 200 python bench.py --host-cores 4 --steps 5 --warmup 3 $LEAN --no-kernel-timing --extra-sync-modes 2,0 \
    > $O/bench_hc4_m1_arena0.json 2> $O/bench_hc4_m1_arena0.err
stamp c4
timeout 400 python bench.py --config c4 --steps 3 --warmup 2 --no-exact > $O/bench_c4.json 2> $O/bench_c4.err; echo "rc=$?" >> $O/bench_c4.err
stamp c3
timeout 600 python bench.py --config c3 --steps 2 --warmup 1 --no-exact --no-cpu-baseline > $O/bench_c3_shard0.json 2> $O/bench_c3.err; echo "rc=$?" >> $O/bench_c3.err
stamp c3hic
timeout 600 python bench.py --config c3hic --steps 2 --warmup 1 --no-exact --no-cpu-baseline > $O/bench_c3hic_shard0.json 2> $O/bench_c3hic.err; echo "rc=$?" >> $O/bench_c3hic.err
stamp ncu
export CL2_NO_POOL=1
python profiles/prof_target.py 8.06e6 1 > $O/prof_target_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_xorder_from_grid|k_os_pass' -c 9 \
    -f -o $O/prof_target_full python profiles/prof_target.py 8.06e6 1 > $O/prof_target_full.log 2>&1
stamp done
