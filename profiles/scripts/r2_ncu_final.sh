# Final round-2 build: (1) launch list of the timed steps of the default workload (three passes over 100 000 candidates),
# (2) launch list of the legs either side of the hot path (from-text path with the device tokenizer, candidate scan, hard filters),
# (3) --set full of the text kernels (tokenizer passes) on the bench workload.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --ncu --steps 2 --warmup 1 --no-cpu-baseline --no-text --no-cli --no-e2e --no-scan --no-filters > gpurun_out/r2f_ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tok_|window_table|count_newlines|write_row_offsets|scan_tiles|scan_kernel|hard_filter|encode_pileup|DeviceScan' -c 300 --csv --log-file gpurun_out/r2f_launches_sides.csv \
    python bench.py --ncu --steps 1 --warmup 1 --no-cpu-baseline --no-cli > gpurun_out/r2f_ncu_sides.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'tok_rows_kernel' -c 2 -s 4 -o gpurun_out/r2f_tok \
    python profiles/debug_device_tokenizer.py > gpurun_out/r2f_ncu_tok.log 2>&1
ncu -i gpurun_out/r2f_tok.ncu-rep --page details > gpurun_out/r2f_tok_details.txt 2>/dev/null
ls -la gpurun_out | tail -8
