# Round profile job (run through gpurun): bench lines, launch list + DRAM traffic, ncu full captures, phase counters.
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_search|k_raster|k_count_cols|k_fill_cells" --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_search_pipe --launch-count 1 -f -o gpurun_out/ksearch_pipe python tools/profile_run.py 1024 4000 > gpurun_out/prof_pipe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_search<" --launch-count 1 -f -o gpurun_out/ksearch_narrow python tools/profile_run.py 1024 4000 > gpurun_out/prof_narrow.log 2>&1
AVP_TRACE_POP=5000 python tools/gpu_pipe_profile.py > gpurun_out/pipe_phase.log 2>&1
python tools/bench_corridor.py > gpurun_out/bench_corridor.json 2> gpurun_out/bench_corridor.err
ncu --set full --clock-control none -k regex:k_corridor --launch-skip 3 --launch-count 1 -f -o gpurun_out/kcorridor python tools/bench_corridor.py 19 262144 > gpurun_out/prof_corridor.log 2>&1
tail -c 400 gpurun_out/bench_n1.json; cat gpurun_out/bench_corridor.json; ls -la gpurun_out
