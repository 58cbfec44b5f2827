# Round profile job (run through gpurun): bench lines, launch list + DRAM traffic, ncu full capture of the pipelined kernel,
# phase counters.  TAG names the output files (gpurun_out/TAG_*).
mkdir -p gpurun_out
T=${1:-prof}
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref_n1.json 2> gpurun_out/${T}_bench_ref_n1.err; echo "ref rc $?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_search|k_raster|k_count_cols|k_fill_cells" --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1; echo "launch list rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_search_pipe --launch-count 1 -f -o gpurun_out/${T}_ksearch_pipe python tools/profile_run.py 1024 4000 > gpurun_out/${T}_prof_pipe.log 2>&1; echo "ncu pipe rc $?"
AVP_TRACE_POP=5000 timeout 200 python tools/gpu_pipe_profile.py > gpurun_out/${T}_pipe_phase.log 2>&1; echo "phase rc $?"
tail -c 600 gpurun_out/${T}_bench_n1.json; ls -la gpurun_out | tail -8
