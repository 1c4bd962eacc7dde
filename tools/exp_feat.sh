#!/bin/bash
# Times every experimental feature-kernel build (lib/exp/*.so) against the default one; with NCU=1 also records their DRAM
# bytes, instruction count and issue rate (one ncu pass per build).
shopt -s nullglob
O=gpurun_out/exp
mkdir -p $O
N=${1:-65536}
for lib in default cough_detector_b200/lib/exp/*.so; do
  if [ $lib = default ]; then unset CD_LIB; else export CD_LIB=$PWD/$lib; fi
  python tools/feat_time.py $N 2>&1 | grep -v "^Trace\|^  \|^Attr" >> $O/times.txt
done
cat $O/times.txt
[ "$NCU" = 1 ] || exit 0
for lib in default cough_detector_b200/lib/exp/*.so; do
  tag=$(basename $lib .so)
  if [ $lib = default ]; then unset CD_LIB; else export CD_LIB=$PWD/$lib; fi
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:feat_kernel -s 1 -c 1 --csv \
      --log-file $O/dram_$tag.csv python tools/feat_profile_case.py $N > $O/ncu_$tag.log 2>&1
  echo "$tag: $(tail -6 $O/dram_$tag.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tr '\n' ' ')"
done
