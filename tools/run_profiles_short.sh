#!/bin/bash
# Trimmed tools/run_profiles.sh for a nearly spent GPU budget: suite, both bench arms, the two launch lists and the --set full
# capture of one slab's seven kernels.  Lands in gpurun_out/f like the long script; tools/collect_profiles.py condenses it.
set -x
O=gpurun_out/f
rm -rf $O; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo smoke=$?
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo pytest=$?; tail -3 $O/pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo ref=$?
python bench.py > $O/bench.json 2> $O/bench.err; echo bench=$?
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"feat|stem|conv3x3|head|map_from|predict|stream|block_max" -c 600 --csv \
    --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --min-seconds 0.01 --no-cpu-baseline --no-sweep > $O/ncu0.log 2>&1; echo launches_bench=$?
python tools/profile_case.py 16384 > $O/pc.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"feat|stem|conv3x3|head" --csv \
      --log-file $O/launches.csv python tools/profile_case.py 16384 > $O/ncu1.log 2>&1; echo launches=$?
ncu --set full --clock-control none --import-source on -k regex:"feat|stem|conv3x3|head" -s 7 -c 7 \
    -o $O/prof_full -f python tools/profile_case.py 16384 > $O/ncu2.log 2>&1; echo full=$?
