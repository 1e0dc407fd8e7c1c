cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aff_layers_kernel -c 2 -o gpurun_out/r2_aff_fused python profiles/phase_timing_aff.py 20000 > gpurun_out/r2_ncu_aff.log 2>&1
ls -la gpurun_out/*.ncu-rep
