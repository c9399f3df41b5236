#!/bin/bash
# Third gpurun call of the round: the final native code (arena blocks confined to one chunk, one blocking-sync event per
# stream) through the GPU tests, smoke, the default bench line and the three other BASELINE configs.  -> gpurun_out/c/
O=gpurun_out/c
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/timeline.log; }
stamp c3hic
timeout 600 python bench.py --config c3hic --steps 2 --warmup 1 --no-exact --no-cpu-baseline > $O/bench_c3hic_shard0.json 2> $O/bench_c3hic.err; echo "rc=$?" >> $O/bench_c3hic.err
tail -2 $O/bench_c3hic.err
stamp tests
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
stamp smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
stamp bench_default
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "rc=$?" >> $O/bench_c2.err
stamp hostcores
CL2_BLOCKING_SYNC=1 timeout 200 python bench.py --host-cores 4 --steps 8 --warmup 3 --no-cpu-baseline --no-exact --no-files --no-kernel-timing \
    --extra-sync-modes 2,0 > $O/bench_hc4_m1.json 2> $O/bench_hc4_m1.err
stamp c3
timeout 600 python bench.py --config c3 --steps 2 --warmup 1 --no-exact --no-cpu-baseline > $O/bench_c3_shard0.json 2> $O/bench_c3.err; echo "rc=$?" >> $O/bench_c3.err
stamp c4
timeout 400 python bench.py --config c4 --steps 3 --warmup 2 --no-exact > $O/bench_c4.json 2> $O/bench_c4.err; echo "rc=$?" >> $O/bench_c4.err
stamp done
