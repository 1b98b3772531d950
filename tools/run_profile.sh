# Round profile run (one B200, under gpurun): plain bench, reference arm, ncu launch list, ncu full capture.
#   bash tools/run_profile.sh <tag>     -> gpurun_out/*_<tag>.*
tag=${1:-r1j}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_pipe_vec3|k_decode_vec3" -c 4 -o gpurun_out/full_$tag -f \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
