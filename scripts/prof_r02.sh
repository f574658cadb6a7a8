#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): launch list of the bench command + `ncu --set full`
# captures of the dominant kernels of every workload; read here with scripts/summarize_ncu.py / ncu_multi.py.
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_bench_under_ncu.json 2> $O/r02_bench_under_ncu.err
$NCU -k regex:dft_kernel -c 1 -o $O/r02_dft_fp32 -f python scripts/prof_dft.py C3 0 1 > $O/r02_prof_dft_fp32.log 2>&1
$NCU -k regex:dft_tc5_kernel -c 1 -o $O/r02_dft_tc5 -f python scripts/prof_dft.py C3 200 1 > $O/r02_prof_dft_tc5.log 2>&1
# one whole fast-mode grid() step (second repetition): hist, scans, work list, record pass, tile kernel, normalise
$NCU -k regex:'gf_|grid_normalise' -s 7 -c 7 -o $O/r02_grid -f python scripts/prof_grid.py 2 > $O/r02_prof_grid.log 2>&1
# one whole galario-path likelihood (third repetition): plane flags, row pass, column pass, fused sampler + chi^2
$NCU -k regex:'plane_nonzero|rfft_|fft_chi2' -s 8 -c 4 -o $O/r02_fft -f python scripts/prof_fft.py > $O/r02_prof_fft.log 2>&1
# one whole NUFFT-path likelihood (third repetition), fused entry point
$NCU -k regex:'plane_nonzero|rfft_|nufft_chi2' -s 8 -c 4 -o $O/r02_nufft -f python scripts/prof_nufft.py C3 > $O/r02_prof_nufft.log 2>&1
ls -la $O | tail -20
tail -n 3 $O/r02_prof_grid.log; tail -n 3 $O/r02_prof_fft.log; tail -n 3 $O/r02_prof_nufft.log
