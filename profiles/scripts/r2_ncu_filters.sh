# ncu --set full of the per-site hard-filter kernel on the bench workload (1600 sites per launch, with source).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hard_filter_kernel' -c 1 -s 2 \
    -o gpurun_out/r2_filters python bench.py --ncu --steps 1 --warmup 1 --no-cpu-baseline --no-text --no-cli --no-e2e --no-scan > gpurun_out/r2_ncu_filters.log 2>&1
ncu -i gpurun_out/r2_filters.ncu-rep --page details > gpurun_out/r2_filters_details.txt 2>/dev/null
ncu -i gpurun_out/r2_filters.ncu-rep --page source --csv > gpurun_out/r2_filters_source.csv 2>/dev/null
ls -la gpurun_out | tail -6
