mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_search|k_raster|k_count_cols|k_fill_cells" --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_search --launch-skip 1 --launch-count 1 -f -o gpurun_out/ksearch_wide_r01 python tools/profile_run.py 384 2000 > gpurun_out/prof_wide.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_search --launch-skip 0 --launch-count 1 -f -o gpurun_out/ksearch_narrow_r01 python tools/profile_run.py 384 2000 > gpurun_out/prof_narrow.log 2>&1
tail -c 600 gpurun_out/bench_n1.json; ls -la gpurun_out
