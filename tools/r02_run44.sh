mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_row_fwd|k_col_fwd|k_lincomb|k_modup2|k_dot|k_tensor" -s 30 -c 16 -o gpurun_out/r02f_generic -f python bench.py --workload eval_relu --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/r02f_generic.log 2>&1
tail -1 gpurun_out/r02f_generic.log | cut -c1-200
ls -la gpurun_out/r02f_generic.ncu-rep
