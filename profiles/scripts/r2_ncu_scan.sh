# ncu --set full of the STEP-1 candidate scan kernel on the bench workload (one launch, with source).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel|write_row_offsets|count_newlines' -c 3 \
    -o gpurun_out/r2_scan python bench.py --ncu --steps 1 --warmup 1 --no-cpu-baseline --no-text --no-cli --no-e2e > gpurun_out/r2_ncu_scan.log 2>&1
ncu -i gpurun_out/r2_scan.ncu-rep --page raw --csv > gpurun_out/r2_scan_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_scan.ncu-rep --page details > gpurun_out/r2_scan_details.txt 2>/dev/null
ls -la gpurun_out | tail -8
