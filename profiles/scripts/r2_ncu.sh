# Round-2 ncu passes (run under gpurun, 1 GPU): (1) launch list of two bench steps, (2) --set full of the heavy kernels of ONE
# engine chunk (37 888 candidates).  The full report stays on the box (> 64 MiB); its raw page is exported as csv, plus small
# reports with source for the GRU and the pair GEMM.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --ncu --steps 2 --warmup 1 --no-cpu-baseline --no-text --no-cli --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
timeout 1500 ncu --set full --clock-control none \
    -k regex:'gru4_kernel|gru1_fused_kernel|gemm_bf16x3_kernel|gemm_pair_kernel|aff_layers_kernel|encode_pileup_kernel|aff_stage1' -c 14 \
    -o /tmp/r2_full python bench.py --ncu --candidates 37888 --steps 1 --warmup 0 --no-cpu-baseline --no-text --no-cli --no-e2e > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gru4_kernel' -c 1 \
    -o gpurun_out/r2_gru4 python profiles/debug_gru4.py 1 > gpurun_out/r2_ncu_gru4.log 2>&1
ls -la gpurun_out/ /tmp/r2_full.ncu-rep | tail -12
