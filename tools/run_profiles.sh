#!/bin/bash
# Runs on the GPU box (gpurun): plain runs first, then the ncu passes of the same commands.  Everything lands in
# gpurun_out/f; tools/collect_profiles.py condenses it into profiles/ on the development machine.
set -x
O=gpurun_out/f
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo smoke=$?
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo pytest=$?; tail -3 $O/pytest.log
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo ref=$?
python bench.py > $O/bench.json 2> $O/bench.err; echo bench=$?
python bench.py --precision fp32 --no-cpu-baseline --no-sweep --no-stream > $O/bench_fp32.json 2> $O/bench_fp32.err; echo fp32=$?
python tools/h2d_matrix.py > $O/h2d_matrix.json 2> $O/h2d_matrix.err; echo h2d=$?
python tools/latency_probe.py > $O/latency.json 2> $O/latency.err; echo latency=$?
# launch list of the bench command itself (it exited 0 plainly just above)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"feat|stem|conv3x3|head|map_from|predict|stream|block_max" -c 600 --csv \
    --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --min-seconds 0.01 --no-cpu-baseline --no-sweep > $O/ncu0.log 2>&1; echo launches_bench=$?
python tools/profile_case.py 16384 > $O/pc.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"feat|stem|conv3x3|head" --csv \
      --log-file $O/launches.csv python tools/profile_case.py 16384 > $O/ncu1.log 2>&1; echo launches=$?
ncu --set full --clock-control none --import-source on -k regex:"feat|stem|conv3x3|head" -s 7 -c 7 \
    -o $O/prof_full -f python tools/profile_case.py 16384 > $O/ncu2.log 2>&1; echo full=$?
python tools/feat_profile_case.py 65536 > $O/fp.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:feat_kernel -s 1 -c 1 \
      -o $O/prof_feat -f python tools/feat_profile_case.py 65536 > $O/ncu3.log 2>&1; echo feat=$?
