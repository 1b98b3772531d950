# Round profile run (one B200, under gpurun): plain bench, reference arm, ncu launch list, ncu full captures.
#   bash tools/run_profile.sh <tag>     -> gpurun_out/*_<tag>.*     (the .ncu-rep files are summarised and deleted: 64 MiB limit)
tag=${1:-r2}
summarise() {   # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --print-source cuda,sass --csv > /tmp/$1_src.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/$1_src.csv 60 > gpurun_out/$1_lines.txt 2>&1
  rm -f gpurun_out/$1.ncu-rep /tmp/$1_src.csv
}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
tools/pcie_probe 0 512 8 > gpurun_out/pcie_$tag.json 2>&1
for c in f32 i64 log; do python tools/prof_groups.py $c 5; done > gpurun_out/groups_$tag.jsonl 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$tag.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_pipe_vec3|k_decode_vec3" -c 4 -o gpurun_out/full_$tag -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_full_$tag.log 2>&1
summarise full_$tag
for c in f32 i64 log; do
  python tools/prof_groups.py $c 3 > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_group_fused -s 1 -c 1 -o gpurun_out/full_group_${c}_$tag -f \
      python tools/prof_groups.py $c 3 > gpurun_out/ncu_group_${c}_$tag.log 2>&1
  summarise full_group_${c}_$tag
done
ls -la gpurun_out | tail -20
